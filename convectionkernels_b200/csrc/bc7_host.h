// Host-side BC7 support: encoding-plan configuration (reference ConvectionKernels_BC67.cpp:3291-3483), the
// compilation of a plan into the kernel's command stream, and the per-launch constant block.
#pragma once

#include <vector>
#include "bc7_core.cuh"

namespace cvttb200
{
    // Kernels::ConfigureBC7EncodingPlanFromFineTuningParams, BC67.cpp:3355-3483.  Always succeeds (returns true) like the reference.
    bool bc7_plan_from_fine_tuning(BC7PlanPOD &plan, const BC7FineTuningPOD &params);
    // Kernels::ConfigureBC7EncodingPlanFromQuality, BC67.cpp:3291-3353
    void bc7_plan_from_quality(BC7PlanPOD &plan, int quality);
    // BC7EncodingPlan::BC7EncodingPlan(), ConvectionKernels.h:166-198
    void bc7_plan_default(BC7PlanPOD &plan);

    // Flattens a plan into the kernel's command stream (see bc7_core.cuh).  Returns the number of result slots used.
    //   kBC7StreamPlain  SHAPE / EVAL / DUAL only (the kernels with per-trial group votes or single-colour candidates)
    //   kBC7StreamPair   PAIR2 commands for the two-subset modes: second subsets as compacted tasks of the CTA
    //   kBC7StreamSplit  small calls: `slices` independent sub-streams behind a table of their offsets (PAIR2 commands too)
    enum { kBC7StreamPlain = 0, kBC7StreamPair = 1, kBC7StreamSplit = 2 };
    int bc7_compile_plan(const BC7PlanPOD &plan, std::vector<uint32_t> &cmds, int form, int slices = 0);

    // Fills everything of BC7Params except `cmds`.  rcpN[n] must hold the host's _mm_rcp_ps((float)n), n = 0..16.
    void bc7_fill_params(BC7Params &P, const OptionsPOD &options, const BC7PlanPOD &plan, const float rcpN[17]);

    const BC7PackTables &bc7_pack_tables();

    // shape tables (tools/gen_tables.py)
    unsigned bc7_shape_mask(int shape);
}
