// TEST INFRASTRUCTURE ONLY.  CPU build of the per-thread decoders (convectionkernels_b200/csrc/decode_core.cuh).  Not part of the
// product library.
#include <string.h>

#include "../../convectionkernels_b200/csrc/bc7_host.h"
#include "../../convectionkernels_b200/csrc/bc6h_host.h"
#include "../../convectionkernels_b200/csrc/decode_core.cuh"

using namespace cvttb200;

// kind: 0 BC7 -> PixelBlockU8, 1 BC6HU, 2 BC6HS -> PixelBlockF16
extern "C" int hostsim_decode(int kind, const uint8_t *encoded, size_t nBlocks, uint8_t *out)
{
    for (size_t b = 0; b < nBlocks; b++)
    {
        uint32_t in[4];
        memcpy(in, encoded + b * 16, 16);
        if (kind == 0)
        {
            uint32_t px[16];
            ArraySink sink = { px };
            bc7_decode_block(bc7_pack_tables(), in, sink);
            memcpy(out + b * 64, px, 64);
        }
        else
        {
            uint32_t px[32];
            ArraySink sink = { px };
            bc6h_decode_block(bc6h_tables(), in, kind == 2, sink);
            memcpy(out + b * 128, px, 128);
        }
    }
    return 0;
}
