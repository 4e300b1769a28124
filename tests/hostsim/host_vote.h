// TEST INFRASTRUCTURE ONLY.  The kernels' 8-lane segment votes, for eight host threads that each run one lane of a
// reference group: a sense-reversing spin barrier with an OR / MAX reduction.
#pragma once
#include <atomic>
#include <thread>

struct GroupShared
{
    std::atomic<int> count{0};
    std::atomic<int> sense{0};
    std::atomic<int> acc[2];
    GroupShared() { acc[0] = 0; acc[1] = 0; }
};

struct HostVote
{
    GroupShared *g;
    int localSense = 0;
    int phase = 0;

    // max over the group of a non-negative value
    int max(int v)
    {
        const int slot = phase & 1;
        phase++;
        localSense ^= 1;
        int cur = g->acc[slot].load();
        while (v > cur && !g->acc[slot].compare_exchange_weak(cur, v)) {}
        if (g->count.fetch_add(1) == 7)
        {
            g->acc[slot ^ 1].store(0);
            g->count.store(0);
            g->sense.store(localSense);
        }
        else
        {
            int spins = 0;
            while (g->sense.load() != localSense)
                if (++spins > 64)
                    std::this_thread::yield();
        }
        return g->acc[slot].load();
    }
    bool any(bool x) { return max(x ? 1 : 0) != 0; }
    bool all(bool x) { return !any(!x); }
    bool warp_any(bool x) { return any(x); } // here the "warp" is the group; callers sit in group-uniform control flow
};
