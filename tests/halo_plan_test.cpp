// tests/halo_plan_test.cpp -- CPU driver of csrc/halo_plan.h for tests/test_halo_plan.py.
//   halo_plan_test W per D chunk merge_gap flags.bin  ->  one line per range: "rank kind peer off count" (kind r = receive, s = send)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../bose-hubbard-phase-transition_b200/csrc/halo_plan.h"

int main(int argc, char** argv)
{
    if (argc < 7) return 2;
    const int W = atoi(argv[1]);
    const long per = atol(argv[2]), D = atol(argv[3]), chunk = atol(argv[4]);
    const int gap = atoi(argv[5]);
    const long nchunks = (per * W + chunk - 1) / chunk;
    std::vector<unsigned char> flags((size_t)W * nchunks);
    FILE* f = fopen(argv[6], "rb");
    if (!f || fread(flags.data(), 1, flags.size(), f) != flags.size()) return 3;
    fclose(f);
    for (int me = 0; me < W; ++me) {
        std::vector<BhHaloRange> recv, send;
        bh_halo_plan(flags.data(), nchunks, W, me, per, D, gap, recv, send, chunk);
        for (const auto& r : recv) printf("%d r %d %ld %ld\n", me, r.peer, (long)r.off, (long)r.count);
        for (const auto& r : send) printf("%d s %d %ld %ld\n", me, r.peer, (long)r.off, (long)r.count);
    }
    return 0;
}
