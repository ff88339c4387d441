// lockstep_sched_test.cpp -- TEST INFRASTRUCTURE ONLY.  Exercises the baton scheduler of the lockstep solves
// (bose-hubbard-phase-transition_b200/csrc/lockstep_sched.h) on the CPU with fake solves: random numbers of host-only steps
// and of shareable requests per point, refill from a list of points.  Checks: never two fibers running at once, every
// request launched exactly once and only while its fiber is parked, groups share one key, no deadlock (watchdog).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <future>
#include <random>
#include <thread>
#include <vector>

#include "../bose-hubbard-phase-transition_b200/csrc/lockstep_sched.h"

struct Sim {
    LockstepSched sched;
    std::atomic<int> owner{-1};
    std::atomic<long> violations{0};
    bool pending[LS_MAX_FIBERS] = {false, false, false, false};
    long launched[LS_MAX_FIBERS] = {0, 0, 0, 0};
    long requested[LS_MAX_FIBERS] = {0, 0, 0, 0};
    long alone[LS_MAX_FIBERS] = {0, 0, 0, 0};
    long group_hist[LS_MAX_FIBERS + 1] = {0, 0, 0, 0, 0};
    int next_point = 0;
};

static int run_case(unsigned seed, int nfib, int npoints, bool mixed_keys)
{
    Sim sim;
    sim.sched.reset(nfib);
    // the work of every point is fixed up front (the same whatever the scheduling)
    std::mt19937 gen(seed);
    std::vector<std::vector<int>> ops(npoints);  // 0 = host-only step, k > 0 = request with key k
    for (auto& o : ops) {
        const int pre = gen() % 6, nreq = 1 + gen() % 40;
        for (int i = 0; i < pre; ++i) o.push_back(mixed_keys ? 1 : 0);
        for (int i = 0; i < nreq; ++i) {
            o.push_back(8);
            if (gen() % 3 == 0) o.push_back(0);
        }
        if (mixed_keys)
            for (int i = 0; i < 3; ++i) o.push_back(1);
    }
    auto launch = [&](const int* grp, int ng) {
        if (ng < 1 || ng > nfib) sim.violations++;
        for (int q = 0; q < ng; ++q) {
            const int f = grp[q];
            if (!sim.sched.parked[f] || !sim.pending[f] || sim.sched.key[f] != sim.sched.key[grp[0]]) sim.violations++;
            sim.pending[f] = false;
            sim.launched[f]++;
        }
        sim.group_hist[ng]++;
        return 0;
    };
    auto body = [&](int me) {
        std::mt19937 g2(seed * 977u + me);
        sim.sched.wait_first_turn(me);
        sim.owner = me;
        for (;;) {
            if (sim.next_point >= npoints) break;
            const int p = sim.next_point++;
            for (int k : ops[p]) {
                if (sim.owner.load() != me) sim.violations++;
                if (g2() % 4 == 0) std::this_thread::yield();
                if (k == 0) continue;
                sim.requested[me]++;
                sim.pending[me] = true;
                int st = 0;
                const bool shared = sim.sched.request(me, k, launch, &st);
                sim.owner = me;
                if (!shared) {
                    sim.pending[me] = false;
                    sim.alone[me]++;
                } else if (sim.pending[me] || st != 0) {
                    sim.violations++;
                }
            }
        }
        if (sim.owner.load() != me) sim.violations++;
        sim.sched.finish(me, launch);
    };
    auto fut = std::async(std::launch::async, [&] {
        std::vector<std::thread> th;
        for (int i = 0; i < nfib; ++i) th.emplace_back(body, i);
        for (auto& t : th) t.join();
    });
    if (fut.wait_for(std::chrono::seconds(20)) != std::future_status::ready) {
        std::fprintf(stderr, "DEADLOCK seed %u nfib %d npoints %d\n", seed, nfib, npoints);
        std::_Exit(2);
    }
    long req = 0, done = 0;
    for (int i = 0; i < nfib; ++i) {
        req += sim.requested[i];
        done += sim.launched[i] + sim.alone[i];
        if (sim.pending[i]) sim.violations++;
    }
    if (req != done || sim.next_point != npoints || sim.sched.turn != -1) sim.violations++;
    if (sim.violations.load()) {
        std::fprintf(stderr, "FAIL seed %u nfib %d npoints %d: %ld violations (requests %ld, served %ld)\n", seed, nfib, npoints,
                     sim.violations.load(), req, done);
        return 1;
    }
    return 0;
}

int main(int argc, char** argv)
{
    const int nseeds = argc > 1 ? std::atoi(argv[1]) : 40;
    int bad = 0, cases = 0;
    for (int seed = 1; seed <= nseeds; ++seed)
        for (int nfib = 2; nfib <= LS_MAX_FIBERS; ++nfib)
            for (int npoints : {nfib, nfib + 1, 3 * nfib + 1})
                for (int mixed = 0; mixed < 2; ++mixed) {
                    bad += run_case((unsigned)seed, nfib, npoints, mixed != 0);
                    ++cases;
                }
    std::printf("%d cases, %d failed\n", cases, bad);
    return bad ? 1 : 0;
}
