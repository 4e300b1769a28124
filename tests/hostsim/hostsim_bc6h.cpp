// TEST INFRASTRUCTURE ONLY.  CPU build of the BC6H per-thread device code (convectionkernels_b200/csrc/bc6h_core.cuh).
// The eight lanes of a reference group run as eight host threads; the group votes of the kernel (ballots over an
// 8-lane segment) become a spin barrier with an OR reduction.  Not part of the product library.
#include <algorithm>
#include <thread>
#include <vector>
#include <xmmintrin.h>

#include "../../convectionkernels_b200/csrc/bc6h_host.h"
#include "host_vote.h"

using namespace cvttb200;

namespace
{
    // lanesPerWarp = 8: one group at a time (warp_any == group any); 32: four groups share the warp-scope votes like the kernel's warps
    template<bool SIGNED, bool FAST>
    void run_lane(int lane, int lanesPerWarp, GroupShared *groupShared, GroupShared *warpShared, const BC6HParams *P, const int16_t *blocks, size_t nBlocks, uint8_t *out)
    {
        HostWarpVote vote;
        vote.group.g = groupShared;
        vote.warp.g = warpShared;
        const BC6HTables &T = bc6h_tables();
        static const int16_t zeroBlock[64] = { 0 };
        for (size_t base = 0; base < nBlocks; base += lanesPerWarp)
        {
            const bool active = base + lane < nBlocks;          // the kernel's ragged last warp encodes zero blocks and drops them
            const int16_t *src = active ? blocks + (base + lane) * 64 : zeroBlock;
            float pw[48];
            uint32_t raw[32], tab[24];
            BC6HLane<1> L;
            L.pw = pw;
            L.raw = raw;
            L.tab = tab;
            for (int px = 0; px < 16; px++)
                bc6h_load_pixel<SIGNED>(*P, L, px, src[px * 4 + 0], src[px * 4 + 1], src[px * 4 + 2]);
            uint32_t o[4];
            bc6h_encode_block<SIGNED, FAST, 1>(*P, T, L, vote, o);
            if (active)
                memcpy(out + (base + lane) * 16, o, 16);
        }
    }
}

static int encode_bc6h(const int16_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options, int isSigned, const float *rcpTable, int lanesPerWarp)
{
    if (nBlocks % 8)
        return -1;
    float rcpN[17];
    for (int n = 0; n < 17; n++)
        rcpN[n] = rcpTable ? rcpTable[n] : _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps((float)n)));
    BC6HParams P;
    bc6h_fill_params(P, *options, rcpN);
    const bool fast = (options->flags & kFlag_BC6H_FastIndexing) != 0;
    GroupShared groups[4], warp;
    warp.size = lanesPerWarp;
    std::vector<std::thread> threads;
    for (int lane = 0; lane < lanesPerWarp; lane++)
    {
        if (isSigned)
            threads.emplace_back(fast ? run_lane<true, true> : run_lane<true, false>, lane, lanesPerWarp, &groups[lane / 8], &warp, &P, blocks, nBlocks, out);
        else
            threads.emplace_back(fast ? run_lane<false, true> : run_lane<false, false>, lane, lanesPerWarp, &groups[lane / 8], &warp, &P, blocks, nBlocks, out);
    }
    for (auto &t : threads)
        t.join();
    return 0;
}

// The small-call launch of bc6h_kernels.cu: the calls of the search dealt out in ranges of callsPerSlice (a CTA each in the
// kernel: here one after the other, each from a fresh best), then per group one re-run of every distinct winner call with the
// lanes' true entry errors.  One group = eight host threads.
namespace
{
    // what a range starts from: the lane's best error after the first seedCalls calls of the search, when they precede the range
    // (the kernel runs those calls in a launch of their own, before the ranges that read them)
    float seed_error(const float *hist, size_t stride, int callBegin, int seedCalls)
    {
        float e = FLT_MAX;
        if (callBegin >= seedCalls)
            for (int c = 0; c < seedCalls; c++)
                e = std::min(e, hist[(size_t)c * stride]);
        return e;
    }

    template<bool SIGNED, bool FAST>
    void run_lane_split(int lane, GroupShared *groupShared, GroupShared *warpShared, const BC6HParams *P, const int16_t *blocks, size_t nBlocks, uint8_t *out, int callsPerSlice, int seedCalls, float *history, int *winners)
    {
        HostWarpVote vote;
        vote.group.g = groupShared;
        vote.warp.g = warpShared;           // its own barrier object, also of eight threads: the warp scope is the group here
        const BC6HTables &T = bc6h_tables();
        for (size_t base = 0; base < nBlocks; base += 8)
        {
            const size_t block = base + lane;
            const int16_t *src = blocks + block * 64;
            float pw[48];
            uint32_t raw[32], tab[24];
            BC6HLane<1> L;
            L.pw = pw;
            L.raw = raw;
            L.tab = tab;
            for (int px = 0; px < 16; px++)
                bc6h_load_pixel<SIGNED>(*P, L, px, src[px * 4 + 0], src[px * 4 + 1], src[px * 4 + 2]);
            float *hist = history + block;          // [call][nBlocks]
            for (int callBegin = 0; callBegin < kBC6HCalls; callBegin += callsPerSlice)
                bc6h_search_calls<SIGNED, FAST, 1>(*P, T, L, vote, callBegin, std::min<int>(kBC6HCalls, callBegin + callsPerSlice), seed_error(hist, nBlocks, callBegin, seedCalls), hist + (size_t)callBegin * nBlocks, nBlocks, true);
            int winner;
            bc6h_history(hist, nBlocks, kBC6HCalls, winner);
            winners[block] = winner;
            vote.any(false);            // barrier: every lane of the group has published its winner
            BC6HBest mine;
            bc6h_best_reset(mine);
            for (int k = 0; k < 8; k++)
            {
                // k-th distinct winner call of the group, in lane order
                int list[8], count = 0;
                for (int j = 0; j < 8; j++)
                {
                    const int w = winners[base + j];
                    bool seen = w < 0;
                    for (int i = 0; i < count; i++)
                        seen = seen || list[i] == w;
                    if (!seen)
                        list[count++] = w;
                }
                if (k >= count)
                    break;
                const int call = list[k];
                int unused;
                BC6HBest best;
                bc6h_best_reset(best);
                best.error = bc6h_history(hist, nBlocks, call, unused);
                bc6h_run_call<SIGNED, FAST, 1>(*P, T, L, vote, call, best);
                if (winner == call)
                    mine = best;
            }
            uint32_t o[4];
            bc6h_pack_block<SIGNED, FAST, 1>(*P, T, L, mine, o);
            memcpy(out + block * 16, o, 16);
            vote.any(false);            // nobody overwrites winners[] of the next group's slots early (distinct slots anyway)
        }
    }
}

extern "C" int hostsim_encode_bc6h_split(const int16_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options, int isSigned, const float *rcpTable, int callsPerSlice, int seedCalls)
{
    if (nBlocks % 8 || callsPerSlice < 1)
        return -1;
    float rcpN[17];
    for (int n = 0; n < 17; n++)
        rcpN[n] = rcpTable ? rcpTable[n] : _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps((float)n)));
    BC6HParams P;
    bc6h_fill_params(P, *options, rcpN);
    const bool fast = (options->flags & kFlag_BC6H_FastIndexing) != 0;
    GroupShared group, warp;
    std::vector<float> history((size_t)kBC6HCalls * nBlocks);
    std::vector<int> winners(nBlocks);
    std::vector<std::thread> threads;
    for (int lane = 0; lane < 8; lane++)
    {
        if (isSigned)
            threads.emplace_back(fast ? run_lane_split<true, true> : run_lane_split<true, false>, lane, &group, &warp, &P, blocks, nBlocks, out, callsPerSlice, seedCalls, history.data(), winners.data());
        else
            threads.emplace_back(fast ? run_lane_split<false, true> : run_lane_split<false, false>, lane, &group, &warp, &P, blocks, nBlocks, out, callsPerSlice, seedCalls, history.data(), winners.data());
    }
    for (auto &t : threads)
        t.join();
    return 0;
}

extern "C" int hostsim_encode_bc6h(const int16_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options, int isSigned, const float *rcpTable)
{
    return encode_bc6h(blocks, nBlocks, out, options, isSigned, rcpTable, 8);
}

// four groups per simulated warp, like the kernel (the exact pruning votes over the warp)
extern "C" int hostsim_encode_bc6h_warp(const int16_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options, int isSigned, const float *rcpTable)
{
    return encode_bc6h(blocks, nBlocks, out, options, isSigned, rcpTable, 32);
}

// number of inputs (0..31743, every precision the modes use) for which the integer form of the endpoint quantiser differs from
// the fp32 directed-rounding form the reference executes
extern "C" int hostsim_bc6h_quantizer_mismatches(void)
{
    int bad = 0;
    const int precisions[] = { 6, 7, 8, 9, 10, 11, 12, 16 };
    for (int e = 0; e <= 31743; e++)
        for (int p : precisions)
        {
            bad += bc6h_quantize_element<false>(e, p) != bc6h_quantize_element_reference<false>(e, p);
            bad += bc6h_quantize_element<true>(e, p) != bc6h_quantize_element_reference<true>(e, p);
            bad += bc6h_quantize_element<true>(-e, p) != bc6h_quantize_element_reference<true>(-e, p);
        }
    return bad;
}

// number of 16-bit patterns with exponent below 31 for which half_bits_to_float (the conversion the kernels use: IEEE value, halved
// for exponent 0) differs from the reference's TwosCLHalfToFloat bit arithmetic
extern "C" int hostsim_bc6h_half_conversion_mismatches(void)
{
    int bad = 0;
    for (uint32_t u = 0; u < 65536; u++)
        if ((u & 0x7c00u) != 0x7c00u)
            bad += !(half_bits_to_float(u) == twoscl_half_to_float((int)u));
    return bad;
}
