// TEST INFRASTRUCTURE ONLY.  CPU build of the ETC per-thread device code (convectionkernels_b200/csrc/etc_core.cuh); eight host
// threads per reference group, see host_vote.h.  Not part of the product library.
#include <thread>
#include <vector>

#include "../../convectionkernels_b200/csrc/etc_host.h"
#include "host_vote.h"

using namespace cvttb200;

namespace
{
    // kind: 0 ETC1, 1 ETC2 RGB, 2 ETC2 RGBA, 3 ETC2 alpha, 4 EAC R11 unsigned, 5 EAC R11 signed, 6 ETC2 RGB + punch-through alpha
    void run_lane(int lane, GroupShared *shared, const ETCParams *P, int kind, const uint8_t *blocks, size_t nBlocks, uint8_t *out)
    {
        HostVote vote;
        vote.g = shared;
        const ETCTables &T = etc_tables();
        const bool uniform = (P->flags & kFlag_Uniform) != 0, bt709 = (P->flags & kFlag_ETC_UseFakeBT709) != 0;
        std::vector<char> scratchMem(etc_scratch_bytes(1));
        ETCScratch S;
        etc_scratch_layout(S, scratchMem.data(), 1);
        for (size_t base = 0; base < nBlocks; base += 8)
        {
            const size_t b = base + lane;
            uint32_t color[2] = { 0, 0 }, alpha[2] = { 0, 0 };
            if (kind <= 3 || kind == 6)
            {
                const uint8_t *src = blocks + b * 64;
                F4 pw[16];
                int a[16];
                uint32_t transparentMask = 0;
                for (int px = 0; px < 16; px++)
                {
                    const uint8_t *s = src + px * 4;
                    if (bt709)
                    {
                        float yuv[3];
                        etc_to_bt709((float)s[0], (float)s[1], (float)s[2], yuv);
                        pw[px].x = yuv[0]; pw[px].y = yuv[1]; pw[px].z = yuv[2];
                    }
                    else
                    {
                        pw[px].x = uniform ? (float)s[0] : (float)s[0] * P->w[0];
                        pw[px].y = uniform ? (float)s[1] : (float)s[1] * P->w[1];
                        pw[px].z = uniform ? (float)s[2] : (float)s[2] * P->w[2];
                    }
                    pw[px].w = as_float((uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24));
                    a[px] = s[3];
                    if (kind == 6 && a[px] < P->punchThreshold)
                    {
                        // CompressETC2Block, ETC.cpp:1705-1718
                        transparentMask |= 1u << px;
                        pw[px].x = pw[px].y = pw[px].z = 0.0f;
                        pw[px].w = as_float((uint32_t)s[3] << 24);
                    }
                }
                ETCLane<1> L;
                L.pw = pw;
                if (kind == 0)
                {
                    if (bt709)
                    {
                        if (uniform) etc1_encode_block<true, true, 1>(*P, T, L, S, color); else etc1_encode_block<false, true, 1>(*P, T, L, S, color);
                    }
                    else if (uniform) etc1_encode_block<true, false, 1>(*P, T, L, S, color); else etc1_encode_block<false, false, 1>(*P, T, L, S, color);
                }
                else if (kind == 1 || kind == 2)
                {
                    if (bt709)
                    {
                        if (uniform) etc2_encode_block<true, true, 1>(*P, T, L, S, vote, color); else etc2_encode_block<false, true, 1>(*P, T, L, S, vote, color);
                    }
                    else if (uniform) etc2_encode_block<true, false, 1>(*P, T, L, S, vote, color); else etc2_encode_block<false, false, 1>(*P, T, L, S, vote, color);
                }
                else if (kind == 6)
                {
                    if (bt709)
                    {
                        if (uniform) etc2_punchthrough_encode_block<true, true, 1>(*P, T, L, S, vote, transparentMask, color); else etc2_punchthrough_encode_block<false, true, 1>(*P, T, L, S, vote, transparentMask, color);
                    }
                    else if (uniform) etc2_punchthrough_encode_block<true, false, 1>(*P, T, L, S, vote, transparentMask, color); else etc2_punchthrough_encode_block<false, false, 1>(*P, T, L, S, vote, transparentMask, color);
                }
                if (kind == 2 || kind == 3)
                    etc_alpha_encode_block(T, a, false, false, alpha);
            }
            else
            {
                const int16_t *src = reinterpret_cast<const int16_t *>(blocks) + b * 16;
                int a[16];
                for (int px = 0; px < 16; px++)
                {
                    // CompressEACBlock, ETC.cpp:2087-2110
                    int v = src[px];
                    if (kind == 5)
                        v = (v < 1023 ? v : 1023) + 1024, v = (v > 1 ? v : 1);
                    else
                        v = (v < 2047 ? v : 2047), v = (v > 0 ? v : 0);
                    a[px] = v;
                }
                etc_alpha_encode_block(T, a, true, kind == 5, alpha);
            }
            uint32_t words[4];
            int n = 0;
            if (kind >= 2 && kind != 6)
            {
                words[n++] = etc_bswap(alpha[0]);
                words[n++] = etc_bswap(alpha[1]);
            }
            if (kind <= 2 || kind == 6)
            {
                words[n++] = etc_bswap(color[0]);
                words[n++] = etc_bswap(color[1]);
            }
            memcpy(out + b * (size_t)(n * 4), words, n * 4);
        }
    }
}

extern "C" int hostsim_encode_etc_alloc(int kind, const uint8_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options, const OptionsPOD *allocOptions);

extern "C" int hostsim_encode_etc(int kind, const uint8_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options)
{
    return hostsim_encode_etc_alloc(kind, blocks, nBlocks, out, options, options);
}

// allocOptions: what the caller's AllocETC2Data received (chroma side axes)
extern "C" int hostsim_encode_etc_alloc(int kind, const uint8_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options, const OptionsPOD *allocOptions)
{
    if (nBlocks % 8)
        return -1;
    ETCParams P;
    etc_fill_params(P, *options, *allocOptions);
    GroupShared shared;
    std::vector<std::thread> threads;
    for (int lane = 0; lane < 8; lane++)
        threads.emplace_back(run_lane, lane, &shared, &P, kind, blocks, nBlocks, out);
    for (auto &t : threads)
        t.join();
    return 0;
}
