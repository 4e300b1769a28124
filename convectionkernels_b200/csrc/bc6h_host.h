// Host-side BC6H support: per-launch constants and the constant tables of the kernel.
#pragma once

#include "bc6h_core.cuh"

namespace cvttb200
{
    // options.seedPoints / refineRoundsBC6H are clamped like BC6HComputer::Pack does (BC67.cpp:2667-2675).
    void bc6h_fill_params(BC6HParams &P, const OptionsPOD &options, const float rcpN[17]);
    const BC6HTables &bc6h_tables();
}
