// TEST INFRASTRUCTURE ONLY.  CPU build of the BC6H per-thread device code (convectionkernels_b200/csrc/bc6h_core.cuh).
// The eight lanes of a reference group run as eight host threads; the group votes of the kernel (ballots over an
// 8-lane segment) become a spin barrier with an OR reduction.  Not part of the product library.
#include <thread>
#include <vector>
#include <xmmintrin.h>

#include "../../convectionkernels_b200/csrc/bc6h_host.h"
#include "host_vote.h"

using namespace cvttb200;

namespace
{
    template<bool SIGNED, bool FAST>
    void run_lane(int lane, GroupShared *shared, const BC6HParams *P, const int16_t *blocks, size_t nBlocks, uint8_t *out)
    {
        HostVote vote;
        vote.g = shared;
        const BC6HTables &T = bc6h_tables();
        for (size_t base = 0; base < nBlocks; base += 8)
        {
            const int16_t *src = blocks + (base + lane) * 64;
            float lin[48], pw[48];
            uint32_t pix[32];
            BC6HLane<1, FAST> L;
            L.lin = lin;
            L.pw = pw;
            L.pix = pix;
            for (int px = 0; px < 16; px++)
                bc6h_load_pixel<SIGNED>(*P, L, px, src[px * 4 + 0], src[px * 4 + 1], src[px * 4 + 2]);
            uint32_t o[4];
            bc6h_encode_block<SIGNED, FAST, 1>(*P, T, L, vote, o);
            memcpy(out + (base + lane) * 16, o, 16);
        }
    }
}

extern "C" int hostsim_encode_bc6h(const int16_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options, int isSigned, const float *rcpTable)
{
    if (nBlocks % 8)
        return -1;
    float rcpN[17];
    for (int n = 0; n < 17; n++)
        rcpN[n] = rcpTable ? rcpTable[n] : _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps((float)n)));
    BC6HParams P;
    bc6h_fill_params(P, *options, rcpN);
    const bool fast = (options->flags & kFlag_BC6H_FastIndexing) != 0;
    GroupShared shared;
    std::vector<std::thread> threads;
    for (int lane = 0; lane < 8; lane++)
    {
        if (isSigned)
            threads.emplace_back(fast ? run_lane<true, true> : run_lane<true, false>, lane, &shared, &P, blocks, nBlocks, out);
        else
            threads.emplace_back(fast ? run_lane<false, true> : run_lane<false, false>, lane, &shared, &P, blocks, nBlocks, out);
    }
    for (auto &t : threads)
        t.join();
    return 0;
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
