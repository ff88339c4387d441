// hv_split_tables.h -- tables of the two-half ("split") matrix-free H.v kernel for chains (hv_split.cu).
//
// The chain is cut after site p-1: a Fock state is (P, S) with P = occupations of sites 0..p-1 and S = occupations
// of sites p..m-1.  In the descending-lexicographic order of the reference's enumeration
// (src/hamiltonian.cpp:60-85) all states with the same P are contiguous, so
//
//     rank(P, S) = off(P) + sufrank(S),   off(P) = sum_{q<p} f[q][R_q],  sufrank(S) = sum_{q>=p} f[q][R_q]
//
// and the vector restricted to the sector "R bosons in the suffix" is a matrix X_R[P][S] whose rows are contiguous.
// The hopping term of src/hamiltonian.cpp:170-191 then splits into
//     (prefix bonds) (x) 1   : y[P][S] += A(P,P') x[P'][S]   -- the same coefficient for a whole row, coalesced reads
//     1 (x) (suffix bonds)   : y[P][S] += B(S,S') x[P][S']   -- one small table per sector, shared by every prefix
//     cut bond (p-1,p) and periodic bond (m-1,0): one product term per direction, sector R <-> R+-1
// and the diagonal of :194-232 is dU(P) + dU(S).  No packed state and no per-row rank arithmetic is needed.
//
// This header is host-only C++ (no CUDA): hv_split.cu uploads the tables, tests/split_emul.cpp replays the
// kernel's loops on them against the oracle.
#pragma once

#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <vector>

#define BH_SPLIT_MAX_SECTORS 16  // n + 1 <= 16

struct SplitSector {
    uint32_t nS;          // suffix configurations with R bosons on s sites
    uint32_t nSpad;       // rounded up to a multiple of 32 (table stride)
    uint32_t nP;          // prefix configurations with n - R bosons on p sites
    uint32_t pfirst;      // first prefix record of the sector (records are sorted by sector, then LEX)
    uint32_t sbase;       // first entry of the sector in the per-suffix arrays
    uint32_t nx, ny;      // warps of a CTA across suffix chunks / across prefix groups (nx * ny = 8)
    uint32_t ncb, ngb;    // chunk blocks and group blocks; the sector has ngb * ncb CTA items
    uint32_t item_first;  // first CTA item of the sector
    uint32_t pad0, pad1;
};

struct SplitNbr {  // one prefix-internal hop of a prefix: row P' and the amplitude sqrt((n_dst + 1) n_src)
    uint32_t off;
    uint32_t code;  // (n_dst + 1) * n_src
    double amp;
};

struct SplitPrefixHdr {  // 32 bytes; the four partner offsets form one 16-byte word
    uint32_t off_cu;  // off(P + e_{p-1}): partner of "boson moves p -> p-1"   (sector R - 1)
    uint32_t off_cd;  // off(P - e_{p-1}): partner of "boson moves p-1 -> p"   (sector R + 1)
    uint32_t off_wu;  // off(P + e_0):     partner of "boson moves m-1 -> 0"   (sector R - 1)
    uint32_t off_wd;  // off(P - e_0):     partner of "boson moves 0 -> m-1"   (sector R + 1)
    uint32_t off;     // LEX rank of (P, first suffix)
    uint32_t info;    // n_0 | n_{p-1} << 4 | cnt << 8 | dU(P) << 16
    uint32_t pad0, pad1;
};

struct SplitTables {
    int m = 0, n = 0, p = 0, s = 0, G = 0, closed = 0;
    int WP = 0;        // prefix-internal hops per record, 2 (p - 1)
    int WS = 0;        // suffix-internal hops per suffix, 2 (s - 1)
    int rec_bytes = 0;  // sizeof(SplitPrefixHdr) + WP * sizeof(SplitNbr)
    uint32_t nitems = 0;
    uint32_t NP = 0, NSpad = 0;
    SplitSector sec[BH_SPLIT_MAX_SECTORS + 1];
    std::vector<unsigned char> prec;  // NP records
    std::vector<uint32_t> sinfo;      // [NSpad] n_p | n_{m-1} << 4 | cnt << 8 | dU(S) << 16
    std::vector<uint32_t> scross;     // [NSpad][4] sufrank of S - e_p (R-1), S + e_p (R+1), S - e_{m-1} (R-1), S + e_{m-1} (R+1)
    std::vector<uint32_t> snbr;       // per sector [WS][nSpad]: sufrank(S') | code << 24 (padding: S itself, code 0)
};

namespace bh_split_detail {

// all occupation vectors of `sites` sites with `total` bosons, descending lexicographic
inline void enumerate(int sites, int total, std::vector<int>& cur, std::vector<std::vector<int>>& out)
{
    if ((int)cur.size() == sites - 1) {
        cur.push_back(total);
        out.push_back(cur);
        cur.pop_back();
        return;
    }
    for (int k = total; k >= 0; --k) {
        cur.push_back(k);
        enumerate(sites, total - k, cur, out);
        cur.pop_back();
    }
}

}  // namespace bh_split_detail

// f[q][R] is BhTables::f (ctx_basis.cu): f[q][R] = [R > 0] C(R - 1 + m - 1 - q, m - 1 - q), row stride fstride ints.
// Builds the tables for the chain of m sites (closed: periodic bond listed), n bosons, p prefix sites, G prefixes per warp.
// max_nx (1, 2, 4, 8): warps of a CTA laid along the suffix direction for long rows; the other 8 / nx cover prefix groups.
inline void bh_split_build(int m, int n, int p, int G, bool closed, const int* f, int fstride, SplitTables& T, int max_nx = 8)
{
    using namespace bh_split_detail;
    if (m < 3 || p < 1 || p > m - 1 || n < 1 || n + 1 > BH_SPLIT_MAX_SECTORS || G < 1 ||
        (max_nx != 1 && max_nx != 2 && max_nx != 4 && max_nx != 8))
        throw std::invalid_argument("bh_split_build");
    const int s = m - p;
    T = SplitTables();
    T.m = m; T.n = n; T.p = p; T.s = s; T.G = G; T.closed = closed ? 1 : 0;
    T.WP = 2 * (p - 1);
    T.WS = 2 * (s - 1);
    T.rec_bytes = (int)sizeof(SplitPrefixHdr) + T.WP * (int)sizeof(SplitNbr);
    auto F = [&](int q, int R) -> int64_t { return (int64_t)f[q * fstride + R]; };
    // off(P): sites 0..p-1 of a state whose remaining R = n - |P| bosons sit on the suffix
    auto off_of = [&](const std::vector<int>& P) -> uint32_t {
        int64_t r = 0;
        int Rq = n;
        for (int q = 0; q < p; ++q) {
            Rq -= P[q];
            r += F(q, Rq);
        }
        return (uint32_t)r;
    };
    // sufrank(S): sites p..m-1 (S[0] = n_p), R = |S|
    auto sufrank_of = [&](const std::vector<int>& S) -> uint32_t {
        int64_t r = 0;
        int Rq = 0;
        for (int i = 0; i < s; ++i) Rq += S[i];
        for (int i = 0; i + 1 < s; ++i) {
            Rq -= S[i];
            r += F(p + i, Rq);
        }
        return (uint32_t)r;
    };

    uint32_t pfirst = 0, sbase = 0, item = 0;
    for (int R = 0; R <= n; ++R) {
        std::vector<std::vector<int>> pre, suf;
        std::vector<int> cur;
        enumerate(p, n - R, cur, pre);
        enumerate(s, R, cur, suf);
        SplitSector& sc = T.sec[R];
        sc.nS = (uint32_t)suf.size();
        sc.nSpad = (sc.nS + 31u) & ~31u;
        sc.nP = (uint32_t)pre.size();
        sc.pfirst = pfirst;
        sc.sbase = sbase;
        const uint32_t nchunks = sc.nSpad / 32;
        sc.nx = nchunks >= 5 ? 8 : nchunks >= 3 ? 4 : nchunks;
        while ((int)sc.nx > max_nx) sc.nx /= 2;
        sc.ny = 8 / sc.nx;
        sc.ncb = (nchunks + sc.nx - 1) / sc.nx;
        const uint32_t ngroups = (sc.nP + G - 1) / G;
        sc.ngb = (ngroups + sc.ny - 1) / sc.ny;
        sc.item_first = item;
        sc.pad0 = sc.pad1 = 0;
        item += sc.ngb * sc.ncb;

        // ---- prefix records ----
        T.prec.resize((size_t)(pfirst + sc.nP) * T.rec_bytes, 0);
        for (uint32_t i = 0; i < sc.nP; ++i) {
            std::vector<int> P = pre[i];
            unsigned char* rec = T.prec.data() + (size_t)(pfirst + i) * T.rec_bytes;
            SplitPrefixHdr* h = (SplitPrefixHdr*)rec;
            SplitNbr* nb = (SplitNbr*)(rec + sizeof(SplitPrefixHdr));
            h->off = off_of(P);
            int cnt = 0, dU = 0;
            for (int q = 0; q < p; ++q) dU += P[q] * (P[q] + 1);
            for (int q = 0; q + 1 < p; ++q) {
                if (P[q + 1] >= 1) {  // boson moves q+1 -> q
                    std::vector<int> Q = P;
                    Q[q]++; Q[q + 1]--;
                    nb[cnt].off = off_of(Q);
                    nb[cnt].code = (uint32_t)((P[q] + 1) * P[q + 1]);
                    nb[cnt].amp = std::sqrt((double)nb[cnt].code);
                    ++cnt;
                }
                if (P[q] >= 1) {  // boson moves q -> q+1
                    std::vector<int> Q = P;
                    Q[q]--; Q[q + 1]++;
                    nb[cnt].off = off_of(Q);
                    nb[cnt].code = (uint32_t)((P[q + 1] + 1) * P[q]);
                    nb[cnt].amp = std::sqrt((double)nb[cnt].code);
                    ++cnt;
                }
            }
            for (int j = cnt; j < T.WP; ++j) {  // padding: own row, zero amplitude
                nb[j].off = h->off;
                nb[j].code = 0;
                nb[j].amp = 0.0;
            }
            h->off_cu = h->off_cd = h->off_wu = h->off_wd = h->off;
            if (R >= 1) {
                std::vector<int> Q = P;
                Q[p - 1]++;
                h->off_cu = off_of(Q);
                Q = P;
                Q[0]++;
                h->off_wu = off_of(Q);
            }
            if (P[p - 1] >= 1) {
                std::vector<int> Q = P;
                Q[p - 1]--;
                h->off_cd = off_of(Q);
            }
            if (P[0] >= 1) {
                std::vector<int> Q = P;
                Q[0]--;
                h->off_wd = off_of(Q);
            }
            h->info = (uint32_t)P[0] | (uint32_t)P[p - 1] << 4 | (uint32_t)cnt << 8 | (uint32_t)dU << 16;
            h->pad0 = h->pad1 = 0;
        }

        // ---- suffix tables ----
        T.sinfo.resize(sbase + sc.nSpad, 0);
        T.scross.resize((size_t)(sbase + sc.nSpad) * 4, 0);
        T.snbr.resize((size_t)(sbase + sc.nSpad) * T.WS, 0);
        uint32_t* nbr = T.snbr.data() + (size_t)sbase * T.WS;  // [WS][nSpad]
        for (uint32_t k = 0; k < sc.nSpad; ++k) {
            const uint32_t self = std::min(k, sc.nS - 1);
            for (int j = 0; j < T.WS; ++j) nbr[(size_t)j * sc.nSpad + k] = self;  // code 0
            if (k >= sc.nS) continue;
            const std::vector<int>& S = suf[k];
            if (sufrank_of(S) != k) throw std::logic_error("bh_split_build: suffix enumeration is not in rank order");
            int cnt = 0, dU = 0;
            for (int i = 0; i < s; ++i) dU += S[i] * (S[i] + 1);
            for (int i = 0; i + 1 < s; ++i) {
                if (S[i + 1] >= 1) {
                    std::vector<int> Q = S;
                    Q[i]++; Q[i + 1]--;
                    nbr[(size_t)cnt * sc.nSpad + k] = sufrank_of(Q) | (uint32_t)((S[i] + 1) * S[i + 1]) << 24;
                    ++cnt;
                }
                if (S[i] >= 1) {
                    std::vector<int> Q = S;
                    Q[i]--; Q[i + 1]++;
                    nbr[(size_t)cnt * sc.nSpad + k] = sufrank_of(Q) | (uint32_t)((S[i + 1] + 1) * S[i]) << 24;
                    ++cnt;
                }
            }
            uint32_t* cr = T.scross.data() + (size_t)(sbase + k) * 4;
            if (S[0] >= 1) { std::vector<int> Q = S; Q[0]--; cr[0] = sufrank_of(Q); }
            if (R + 1 <= n) { std::vector<int> Q = S; Q[0]++; cr[1] = sufrank_of(Q); }
            if (S[s - 1] >= 1) { std::vector<int> Q = S; Q[s - 1]--; cr[2] = sufrank_of(Q); }
            if (R + 1 <= n) { std::vector<int> Q = S; Q[s - 1]++; cr[3] = sufrank_of(Q); }
            T.sinfo[sbase + k] = (uint32_t)S[0] | (uint32_t)S[s - 1] << 4 | (uint32_t)cnt << 8 | (uint32_t)dU << 16;
        }
        pfirst += sc.nP;
        sbase += sc.nSpad;
    }
    for (int R = n + 1; R <= BH_SPLIT_MAX_SECTORS; ++R) {
        T.sec[R] = SplitSector();
        T.sec[R].item_first = item;
    }
    T.nitems = item;
    T.NP = pfirst;
    T.NSpad = sbase;
}
