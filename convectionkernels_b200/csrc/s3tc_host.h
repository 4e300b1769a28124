// Host-side BC1-BC5 support: per-launch constants.
#pragma once

#include "s3tc_core.cuh"

namespace cvttb200
{
    void s3tc_fill_params(S3TCParams &P, const OptionsPOD &options, const float rcpN[17]);
}
