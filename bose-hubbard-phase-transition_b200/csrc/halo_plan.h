// halo_plan.h -- host-only planning of the halo exchange of the row-partitioned H.v (dist.cu); no CUDA, no NCCL, so that the
// logic is tested on the CPU (tests/halo_plan_test.cpp, tests/test_halo_plan.py).
//
// Every rank owns the slice [per * r, min(D, per * (r + 1))) of the LEX rank range.  flags[reader][c] != 0 means: some hop of a
// row of rank `reader` reads an element of chunk c (HALO_CHUNK consecutive global elements) outside reader's own slice.
// From the all-gathered flags every rank derives the same lists: what it receives from each peer and what it sends to each
// peer -- the send list of (p -> r) and the receive list of (r <- p) are produced by the same call with the same arguments,
// so the grouped ncclSend / ncclRecv of the two sides match element for element.
#pragma once

#include <algorithm>
#include <cstdint>
#include <vector>

#define HALO_CHUNK 4096

struct BhHaloRange {
    int peer;
    int64_t off, count;  // off = global element offset
};

// Ranges of `owner`'s slice that `reader` reads: runs of flagged chunks clipped to the owner's slice; runs separated by at
// most `merge_gap` clean chunks are merged (fewer, larger messages).  Appends to out with .peer = peer.
inline void bh_halo_ranges(const unsigned char* flags_of_reader, int64_t per, int64_t D, int owner, int peer, int merge_gap,
                           std::vector<BhHaloRange>& out, int64_t chunk = HALO_CHUNK)
{
    const int64_t lo = per * owner, hi = std::min<int64_t>(D, per * (owner + 1));
    if (hi <= lo) return;
    const int64_t c0 = lo / chunk, c1 = (hi + chunk - 1) / chunk;
    int64_t run0 = -1, last = -1;
    auto flush = [&]() {
        if (run0 < 0) return;
        const int64_t a = std::max(lo, run0 * chunk), b = std::min(hi, (last + 1) * chunk);
        if (b > a) out.push_back({peer, a, b - a});
        run0 = -1;
    };
    for (int64_t c = c0; c < c1; ++c) {
        if (!flags_of_reader[c]) continue;
        if (run0 >= 0 && c - last > merge_gap) flush();
        if (run0 < 0) run0 = c;
        last = c;
    }
    flush();
}

// The two lists of rank `me` out of the flags of all W ranks ([W][nchunks], nchunks = ceil(per * W / chunk)).
inline void bh_halo_plan(const unsigned char* flags, int64_t nchunks, int W, int me, int64_t per, int64_t D, int merge_gap,
                         std::vector<BhHaloRange>& recv, std::vector<BhHaloRange>& send, int64_t chunk = HALO_CHUNK)
{
    recv.clear();
    send.clear();
    for (int p = 0; p < W; ++p) {
        if (p == me) continue;
        bh_halo_ranges(flags + (size_t)me * nchunks, per, D, p, p, merge_gap, recv, chunk);  // what I read from p's slice
        bh_halo_ranges(flags + (size_t)p * nchunks, per, D, me, p, merge_gap, send, chunk);  // what p reads from mine
    }
}
