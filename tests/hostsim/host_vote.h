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
    int size = 8;                       // threads that meet at this barrier
    GroupShared() { acc[0] = 0; acc[1] = 0; }
};

// max over the threads of a GroupShared (a non-negative value), with the barrier that implies
struct HostReduce
{
    GroupShared *g = nullptr;
    int localSense = 0;
    int phase = 0;

    int max(int v)
    {
        const int slot = phase & 1;
        phase++;
        localSense ^= 1;
        int cur = g->acc[slot].load();
        while (v > cur && !g->acc[slot].compare_exchange_weak(cur, v)) {}
        if (g->count.fetch_add(1) == g->size - 1)
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
};

// The kernel's SegmentVote for a simulated warp of four groups (32 host threads): any / all over the lane's group,
// warp_any over all 32 lanes -- the scope the kernel's exact pruning tests vote in.
struct HostWarpVote
{
    HostReduce group, warp;
    int max(int v) { return group.max(v); }
    bool any(bool x) { return group.max(x ? 1 : 0) != 0; }
    bool all(bool x) { return !any(!x); }
    bool warp_any(bool x) { return warp.max(x ? 1 : 0) != 0; }
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
