// split_emul.cpp -- TEST INFRASTRUCTURE ONLY.  Replays, on the CPU, the loops of the split matrix-free H.v kernel
// (bose-hubbard-phase-transition_b200/csrc/hv_split.cu) over the tables of hv_split_tables.h, so that the table
// construction and the work decomposition can be checked against the oracle without a GPU.  Not part of the product.
#include <cstring>
#include <vector>

#include "../bose-hubbard-phase-transition_b200/csrc/hv_split_tables.h"

static long long binom(int a, int b)
{
    if (b < 0 || b > a) return 0;
    long long r = 1;
    for (int i = 1; i <= b; ++i) r = r * (a - b + i) / i;
    return r;
}

extern "C" int split_emul_hv(int m, int n, int p, int G, int closed, double cJ, double cU, double cmu, long D,
                             const double* x, double* y, int* touched)
{
    const int FS = 18;
    std::vector<int> f(16 * FS, 0);
    for (int q = 0; q < m - 1; ++q)
        for (int R = 0; R <= n + 1; ++R) f[q * FS + R] = R > 0 ? (int)binom(R - 1 + m - 1 - q, m - 1 - q) : 0;
    SplitTables T;
    try {
        bh_split_build(m, n, p, G, closed != 0, f.data(), FS, T);
    } catch (const std::exception&) {
        return -1;
    }
    std::vector<double> sq(256);
    for (int a = 0; a < 256; ++a) sq[a] = std::sqrt((double)a);
    std::memset(touched, 0, sizeof(int) * D);
    const double shift = -(double)n * cmu;
    for (uint32_t item = 0; item < T.nitems; ++item) {
        int R = 0;
        while (R < n && item >= T.sec[R + 1].item_first) ++R;
        const SplitSector& sc = T.sec[R];
        const uint32_t local = item - sc.item_first;
        const uint32_t gb = local / sc.ncb, cb = local % sc.ncb;
        for (uint32_t warp = 0; warp < 8; ++warp) {
            const uint32_t wx = warp % sc.nx, wy = warp / sc.nx;
            const uint32_t chunk = cb * sc.nx + wx;
            if (chunk * 32 >= sc.nSpad) continue;
            const int64_t g0 = ((int64_t)gb * sc.ny + wy) * G;
            const int gcount = (int)std::min<int64_t>(G, (int64_t)sc.nP - g0);
            if (gcount <= 0) continue;
            for (uint32_t lane = 0; lane < 32; ++lane) {
                const uint32_t S = chunk * 32 + lane;  // < nSpad
                const bool valid = S < sc.nS;
                const uint32_t Sc = valid ? S : sc.nS - 1;
                const uint32_t si = T.sinfo[sc.sbase + S];
                const int np = si & 15, nl = (si >> 4) & 15, scnt = (si >> 8) & 255, dUs = si >> 16;
                const uint32_t* cr = &T.scross[(size_t)(sc.sbase + S) * 4];
                for (int g = 0; g < gcount; ++g) {
                    const unsigned char* rec = T.prec.data() + (size_t)(sc.pfirst + g0 + g) * T.rec_bytes;
                    const SplitPrefixHdr* h = (const SplitPrefixHdr*)rec;
                    const SplitNbr* nb = (const SplitNbr*)(rec + sizeof(SplitPrefixHdr));
                    const int n0 = h->info & 15, nq = (h->info >> 4) & 15, pcnt = (h->info >> 8) & 255, dUp = h->info >> 16;
                    double acc = 0.0;
                    for (int j = 0; j < T.WS; ++j) {  // kernel: j < warp max of scnt; padding has amplitude 0
                        const uint32_t v = T.snbr[(size_t)sc.sbase * T.WS + (size_t)j * sc.nSpad + S];
                        if (j >= scnt && (v >> 24) != 0) return -2;
                        acc = std::fma(sq[v >> 24], x[h->off + (v & 0xffffffu)], acc);
                    }
                    for (int j = 0; j < pcnt; ++j) acc = std::fma(nb[j].amp, x[nb[j].off + Sc], acc);
                    if (R >= 1) {
                        if (np) acc = std::fma(sq[(nq + 1) * np], x[h->off_cu + cr[0]], acc);
                        if (closed && nl) acc = std::fma(sq[(n0 + 1) * nl], x[h->off_wu + cr[2]], acc);
                    }
                    if (nq) acc = std::fma(sq[(np + 1) * nq], x[h->off_cd + cr[1]], acc);
                    if (closed && n0) acc = std::fma(sq[(nl + 1) * n0], x[h->off_wd + cr[3]], acc);
                    if (!valid) continue;
                    const long l = (long)h->off + S;
                    if (l >= D) return -3;
                    const double diag = (double)(dUp + dUs) * cU + shift;
                    y[l] = diag * x[l] - (2.0 * cJ) * acc;
                    touched[l]++;
                }
            }
        }
    }
    return 0;
}
