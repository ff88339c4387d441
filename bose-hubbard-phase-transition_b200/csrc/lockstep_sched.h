// lockstep_sched.h -- baton scheduling of the lockstep solves of batch.cu (host-only C++, no CUDA; tested on the CPU by
// tests/lockstep_sched_test.cpp).
//
// Up to LS_MAX_FIBERS host threads ("fibers") each run an ordinary single-point solve.  Exactly one fiber runs at any
// time (it "holds the baton", turn == its index); every fiber launches on the same CUDA stream, so the order of GPU work is
// the order in which the baton holder enqueued it and no cross-stream synchronisation is needed.
//   * request(me, key, launch): fiber `me` wants an operator application that can share a launch with requests of the
//     same key (the polynomial degree).  The fiber parks.  If another fiber can still run, the baton goes there and `me`
//     sleeps until its request has been launched AND it is its turn again.  If every live fiber is parked, `me` calls
//     launch(group, n) once per key group (all parked fibers of that key, under the lock), un-parks everybody and keeps
//     the baton.  With fewer than two live fibers nothing is shared: request() returns false and the caller launches alone.
//   * finish(me, launch): the fiber has no more work.  The baton goes to a runnable fiber; if all remaining live fibers
//     are parked their requests are launched first (otherwise nobody would ever launch them).
#pragma once

#include <condition_variable>
#include <mutex>

#define LS_MAX_FIBERS 4

struct LockstepSched {
    std::mutex mu;
    std::condition_variable cv;
    int nfib = 0;
    int turn = -1;  // fiber allowed to run; -1 when everybody has finished
    bool done[LS_MAX_FIBERS] = {false, false, false, false};
    bool parked[LS_MAX_FIBERS] = {false, false, false, false};
    int key[LS_MAX_FIBERS] = {0, 0, 0, 0};
    int rc[LS_MAX_FIBERS] = {0, 0, 0, 0};

    void reset(int n)
    {
        nfib = n;
        turn = 0;
        for (int j = 0; j < LS_MAX_FIBERS; ++j) {
            done[j] = parked[j] = false;
            key[j] = rc[j] = 0;
        }
    }

    // fiber i blocks here until it is given the baton for the first time
    void wait_first_turn(int i)
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return turn == i; });
    }

    int next_runnable(int after) const
    {
        for (int o = 1; o <= nfib; ++o) {
            const int j = (after + o) % nfib;
            if (!done[j] && !parked[j]) return j;
        }
        return -1;
    }

    // with mu held and no runnable fiber: launch every parked request, one call per key group; returns the first parked fiber
    template <class Launch>
    int launch_parked(Launch&& launch)
    {
        int all[LS_MAX_FIBERS], na = 0;
        for (int j = 0; j < nfib; ++j)
            if (!done[j] && parked[j]) all[na++] = j;
        bool taken[LS_MAX_FIBERS] = {false, false, false, false};
        int result = 0;
        for (int a0 = 0; a0 < na; ++a0) {
            if (taken[a0]) continue;
            int grp[LS_MAX_FIBERS], ng = 0;
            for (int a1 = a0; a1 < na; ++a1)
                if (!taken[a1] && key[all[a1]] == key[all[a0]]) {
                    grp[ng++] = all[a1];
                    taken[a1] = true;
                }
            if (result == 0) result = launch(grp, ng);
        }
        for (int j = 0; j < na; ++j) {
            rc[all[j]] = result;
            parked[all[j]] = false;
        }
        return na ? all[0] : -1;
    }

    // returns false: fewer than two live fibers, the caller launches alone; true: launched (status in *status)
    template <class Launch>
    bool request(int me, int k, Launch&& launch, int* status)
    {
        std::unique_lock<std::mutex> lk(mu);
        int live = 0;
        for (int j = 0; j < nfib; ++j) live += done[j] ? 0 : 1;
        if (live < 2) return false;
        key[me] = k;
        rc[me] = 0;
        parked[me] = true;
        const int nxt = next_runnable(me);
        if (nxt >= 0) {
            turn = nxt;
            cv.notify_all();
            cv.wait(lk, [&] { return turn == me && !parked[me]; });
        } else {
            launch_parked(launch);  // keeps the baton
        }
        *status = rc[me];
        return true;
    }

    template <class Launch>
    void finish(int me, Launch&& launch)
    {
        std::unique_lock<std::mutex> lk(mu);
        done[me] = true;
        int nxt = next_runnable(me);
        if (nxt < 0) nxt = launch_parked(launch);
        turn = nxt;
        cv.notify_all();
    }
};
