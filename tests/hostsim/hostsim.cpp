// TEST INFRASTRUCTURE ONLY.  CPU build of the per-thread device code (convectionkernels_b200/csrc/*_core.cuh)
// so the kernel logic can be checked against the oracle in the GPU-less dev container.  It emulates the
// kernel's warp of 32 lanes = 4 reference groups; it is NOT part of the product library and the product never
// falls back to it.
#include <stdlib.h>
#include <thread>
#include <vector>
#include <xmmintrin.h>

#include "../../convectionkernels_b200/csrc/bc7_host.h"
#include "host_vote.h"

using namespace cvttb200;

namespace
{
    template<bool PUNCH, class Vote>
    void encode_lane(const BC7Params &P, const BC7PackTables &T, bool fast, const BC7LaneFlags &lf, Vote &vote, const uint8_t *block, uint8_t *out)
    {
        uint32_t raw[16];
        memcpy(raw, block, 64);
        F4 gv[16], gw[16];
        BC7Lane<1> L;
        L.raw = raw;
        L.gv = gv;
        L.gw = gw;
        uint32_t o[4];
        BC7SoloExchange ex;                 // PAIR2 commands: the lane searches its own second subsets
        ex.raw = raw;
        if (P.splitSlices > 0)
        {
            // the small-call launch: one search per slice (a CTA each in the kernel), winners through the candidate
            // records, then bc7_finish_kernel's reduction
            BC7Work work, best;
            bc7_work_reset(best);
            for (int s = 0; s < P.splitSlices; s++)
            {
                if (fast)
                    bc7_search_block<true, 1, PUNCH>(P, L, lf, vote, ex, P.cmds + P.cmds[s], work);
                else
                    bc7_search_block<false, 1, PUNCH>(P, L, lf, vote, ex, P.cmds + P.cmds[s], work);
                BC7Candidate c;
                bc7_candidate_pack(work, c);
                bc7_candidate_merge(best, c);
            }
            if (fast)
                bc7_finish_block<true, 1>(P, T, L, best, o);
            else
                bc7_finish_block<false, 1>(P, T, L, best, o);
        }
        else if (fast)
            bc7_encode_block<true, 1, PUNCH>(P, T, L, lf, vote, ex, o);
        else
            bc7_encode_block<false, 1, PUNCH>(P, T, L, lf, vote, ex, o);
        memcpy(out, o, 16);
    }
}

extern "C" int hostsim_encode_bc7(const uint8_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options, const BC7PlanPOD *plan, const float *rcpTable, int warpFlagsAllTrue)
{
    if (nBlocks % 8)
        return -1;
    float rcpN[17];
    for (int n = 0; n < 17; n++)
        rcpN[n] = rcpTable ? rcpTable[n] : _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps((float)n)));
    std::vector<uint32_t> cmds;
    // the same choice of command stream as launch_bc7 makes
    const bool pairCommands = (options->flags & (kFlag_BC7_RespectPunchThrough | kFlag_BC7_TrySingleColor)) == 0;
    // CVTT_HOSTSIM_SPLIT=n: the small-call launch with n slices
    const int slices = (pairCommands && getenv("CVTT_HOSTSIM_SPLIT")) ? atoi(getenv("CVTT_HOSTSIM_SPLIT")) : 0;
    int slots = bc7_compile_plan(*plan, cmds, slices > 0 ? kBC7StreamSplit : (pairCommands ? kBC7StreamPair : kBC7StreamPlain), slices);
    if (slots > kBC7MaxSlots)
        return -2;
    BC7Params P;
    bc7_fill_params(P, *options, *plan, rcpN);
    P.cmds = cmds.data();
    P.splitSlices = slices > 0 ? slices : 0;
    const BC7PackTables &T = bc7_pack_tables();
    const bool fast = (options->flags & kFlag_BC7_FastIndexing) != 0;

    for (size_t warpBase = 0; warpBase < nBlocks; warpBase += 32)
    {
        const size_t lanes = std::min<size_t>(32, nBlocks - warpBase);
        BC7LaneFlags lf[32];
        int minAlpha[32], maxAlpha[32];
        bool punch[32];
        for (size_t l = 0; l < lanes; l++)
        {
            minAlpha[l] = 255;
            maxAlpha[l] = 0;
            punch[l] = true;
            for (int px = 0; px < 16; px++)
            {
                const int a = blocks[(warpBase + l) * 64 + px * 4 + 3];
                minAlpha[l] = std::min<int>(minAlpha[l], a);
                maxAlpha[l] = std::max<int>(maxAlpha[l], a);
                punch[l] = punch[l] && (a == 0 || a == 255);
            }
        }
        bool wRGB = false, wPCA4 = false, wM7 = false;
        for (size_t l = 0; l < lanes; l++)
        {
            const size_t g = l & ~(size_t)7;
            bool anyAlpha = false, allowRGB = false;
            for (size_t k = g; k < g + 8; k++)
            {
                anyAlpha |= minAlpha[k] < 255;
                allowRGB |= minAlpha[k] > 250;
            }
            lf[l].anyBlockHasAlpha = anyAlpha;
            lf[l].allowRGBModes = allowRGB;
            lf[l].blockHasNonMaxAlpha = minAlpha[l] < 255;
            lf[l].blockHasNonZeroAlpha = maxAlpha[l] > 0;
            lf[l].isPunchThrough = punch[l];
            wRGB |= allowRGB;
            wPCA4 |= anyAlpha || !allowRGB;
            wM7 |= anyAlpha || plan->mode7RGBPartitionEnabled != 0;
        }
        for (size_t l = 0; l < lanes; l++)
        {
            lf[l].warpAnyRGB = warpFlagsAllTrue ? true : wRGB;
            lf[l].warpAnyPCA4 = warpFlagsAllTrue ? true : wPCA4;
            lf[l].warpAnyExpand = true;
            lf[l].warpHasWork = true;
            lf[l].warpAnyMode7 = warpFlagsAllTrue ? true : wM7;
        }
        if (options->flags & kFlag_BC7_RespectPunchThrough)
        {
            // per-trial group votes: the eight lanes of a group run as eight threads (host_vote.h)
            for (size_t g = 0; g < lanes; g += 8)
            {
                GroupShared shared;
                std::vector<std::thread> threads;
                for (size_t l = g; l < g + 8; l++)
                    threads.emplace_back([&, l]()
                    {
                        HostVote vote;
                        vote.g = &shared;
                        encode_lane<true>(P, T, fast, lf[l], vote, blocks + (warpBase + l) * 64, out + (warpBase + l) * 16);
                    });
                for (auto &t : threads)
                    t.join();
            }
        }
        else
        {
            BC7NoVote vote;
            for (size_t l = 0; l < lanes; l++)
                encode_lane<false>(P, T, fast, lf[l], vote, blocks + (warpBase + l) * 64, out + (warpBase + l) * 16);
        }
    }
    return 0;
}

extern "C" void hostsim_plan_from_quality(BC7PlanPOD *plan, int q) { bc7_plan_from_quality(*plan, q); }
extern "C" void hostsim_plan_default(BC7PlanPOD *plan) { bc7_plan_default(*plan); }
extern "C" int hostsim_plan_from_finetune(BC7PlanPOD *plan, const BC7FineTuningPOD *ft) { return bc7_plan_from_fine_tuning(*plan, *ft) ? 1 : 0; }
