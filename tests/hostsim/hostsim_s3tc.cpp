// TEST INFRASTRUCTURE ONLY.  CPU build of the BC1-BC5 per-thread device code (convectionkernels_b200/csrc/s3tc_core.cuh); eight host
// threads per reference group (the exhaustive search needs the group maximum, see host_vote.h).  Not part of the product library.
#include <algorithm>
#include <thread>
#include <vector>
#include <xmmintrin.h>

#include "../../convectionkernels_b200/csrc/s3tc_host.h"
#include "host_vote.h"

using namespace cvttb200;

namespace
{
    void run_lane(int lane, GroupShared *shared, const S3TCParams *P, int fmt, const uint8_t *blocks, size_t nBlocks, uint8_t *out)
    {
        HostVote vote;
        vote.g = shared;
        const bool isSigned = (fmt == 5 || fmt == 7);
        const size_t outBytes = (fmt == 1 || fmt == 4 || fmt == 5) ? 8 : 16;
        for (size_t base = 0; base < nBlocks; base += 8)
        {
            const size_t b = base + lane;
            F4 px[16];
            for (int i = 0; i < 16; i++)
            {
                float c[4];
                for (int ch = 0; ch < 4; ch++)
                {
                    int v = blocks[b * 64 + i * 4 + ch];
                    if (isSigned)
                        v = std::max<int>((int8_t)v, -127) + 127;       // Util::BiasSignedInput
                    c[ch] = (float)v;
                }
                px[i].x = c[0]; px[i].y = c[1]; px[i].z = c[2]; px[i].w = c[3];
            }
            S3TCLane<1> L;
            L.px = px;
            uint32_t w[4] = { 0, 0, 0, 0 };
            switch (fmt)
            {
            case 1: s3tc_pack_rgb<1>(*P, L, true, vote, w); break;
            case 2: s3tc_pack_explicit_alpha<1>(L, 3, w); s3tc_pack_rgb<1>(*P, L, false, vote, w + 2); break;
            case 3: s3tc_pack_interpolated_alpha<1>(*P, L, 3, false, w); s3tc_pack_rgb<1>(*P, L, false, vote, w + 2); break;
            case 4: case 5: s3tc_pack_interpolated_alpha<1>(*P, L, 0, isSigned, w); break;
            default: s3tc_pack_interpolated_alpha<1>(*P, L, 0, isSigned, w); s3tc_pack_interpolated_alpha<1>(*P, L, 1, isSigned, w + 2); break;
            }
            memcpy(out + b * outBytes, w, outBytes);
        }
    }
}

// fmt: the cvttb200_format ids 1..7 (BC1, BC2, BC3, BC4U, BC4S, BC5U, BC5S)
extern "C" int hostsim_encode_s3tc(int fmt, const uint8_t *blocks, size_t nBlocks, uint8_t *out, const OptionsPOD *options, const float *rcpTable)
{
    if (nBlocks % 8 || fmt < 1 || fmt > 7)
        return -1;
    float rcpN[17];
    for (int n = 0; n < 17; n++)
        rcpN[n] = rcpTable ? rcpTable[n] : _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps((float)n)));
    S3TCParams P;
    s3tc_fill_params(P, *options, rcpN);
    GroupShared shared;
    std::vector<std::thread> threads;
    for (int lane = 0; lane < 8; lane++)
        threads.emplace_back(run_lane, lane, &shared, &P, fmt, blocks, nBlocks, out);
    for (auto &t : threads)
        t.join();
    return 0;
}
