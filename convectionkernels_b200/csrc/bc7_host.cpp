// Host-side BC7 support (see bc7_host.h).
#include "bc7_host.h"

#include <stdlib.h>
#include <algorithm>
#include <string.h>

#include "bc7_tables.inc"

namespace cvttb200
{
    unsigned bc7_shape_mask(int shape) { return kBC7ShapeMask[shape]; }

    static int popcount16(unsigned m)
    {
        int n = 0;
        for (; m; m &= m - 1)
            n++;
        return n;
    }

    // ---------------------------------------------------------------------------------------------------
    // plan configuration

    void bc7_plan_default(BC7PlanPOD &plan)
    {
        memset(&plan, 0, sizeof(plan));
        for (int i = 0; i < 243; i++)
        {
            plan.rgbShapeList[i] = (uint8_t)i;
            plan.seedPointsForShapeRGB[i] = 4;
        }
        plan.rgbNumShapesToEvaluate = 243;
        for (int i = 0; i < 129; i++)
        {
            plan.rgbaShapeList[i] = (uint8_t)i;
            plan.seedPointsForShapeRGBA[i] = 4;
        }
        plan.rgbaNumShapesToEvaluate = 129;
        plan.mode0PartitionEnabled = 0xffff;
        plan.mode1PartitionEnabled = plan.mode2PartitionEnabled = plan.mode3PartitionEnabled = ~(uint64_t)0;
        plan.mode6Enabled = 1;
        plan.mode7RGBPartitionEnabled = plan.mode7RGBAPartitionEnabled = ~(uint64_t)0;
        for (int r = 0; r < 4; r++)
        {
            plan.mode4SP[r][0] = plan.mode4SP[r][1] = 4;
            plan.mode5SP[r] = 4;
        }
    }

    namespace
    {
        struct ShapeSeeds
        {
            BC7PlanPOD &plan;
            void rgb(int shape, uint8_t sp) { plan.seedPointsForShapeRGB[shape] = std::max(plan.seedPointsForShapeRGB[shape], sp); }
            void rgba(int shape, uint8_t sp) { plan.seedPointsForShapeRGBA[shape] = std::max(plan.seedPointsForShapeRGBA[shape], sp); }
        };
    }

    bool bc7_plan_from_fine_tuning(BC7PlanPOD &plan, const BC7FineTuningPOD &params)
    {
        memset(&plan, 0, sizeof(plan));
        ShapeSeeds seeds = { plan };

        // three-subset modes take their shapes from kBC7Shapes3, two-subset modes from kBC7Shapes2
        for (int p = 0; p < 16; p++)
            if (uint8_t sp = params.mode0SP[p])
            {
                plan.mode0PartitionEnabled |= (uint16_t)(1u << p);
                for (int k = 0; k < 3; k++)
                    seeds.rgb(kBC7Shapes3[p * 3 + k], sp);
            }
        for (int p = 0; p < 64; p++)
        {
            const uint64_t bit = (uint64_t)1 << p;
            if (uint8_t sp = params.mode1SP[p])
            {
                plan.mode1PartitionEnabled |= bit;
                seeds.rgb(kBC7Shapes2[p * 2], sp);
                seeds.rgb(kBC7Shapes2[p * 2 + 1], sp);
            }
            if (uint8_t sp = params.mode2SP[p])
            {
                plan.mode2PartitionEnabled |= bit;
                for (int k = 0; k < 3; k++)
                    seeds.rgb(kBC7Shapes3[p * 3 + k], sp);
            }
            if (uint8_t sp = params.mode3SP[p])
            {
                plan.mode3PartitionEnabled |= bit;
                seeds.rgb(kBC7Shapes2[p * 2], sp);
                seeds.rgb(kBC7Shapes2[p * 2 + 1], sp);
            }
            if (uint8_t sp = params.mode7SP[p])
            {
                plan.mode7RGBAPartitionEnabled |= bit;
                seeds.rgba(kBC7Shapes2[p * 2], sp);
                seeds.rgba(kBC7Shapes2[p * 2 + 1], sp);
            }
        }
        memcpy(plan.mode4SP, params.mode4SP, sizeof(plan.mode4SP));
        memcpy(plan.mode5SP, params.mode5SP, sizeof(plan.mode5SP));
        if (params.mode6SP)
        {
            plan.mode6Enabled = 1;
            seeds.rgba(0, params.mode6SP);
        }

        for (int s = 0; s < 243; s++)
            if (plan.seedPointsForShapeRGB[s])
                plan.rgbShapeList[plan.rgbNumShapesToEvaluate++] = (uint8_t)s;
        for (int s = 0; s < 129; s++)
            if (plan.seedPointsForShapeRGBA[s])
                plan.rgbaShapeList[plan.rgbaNumShapesToEvaluate++] = (uint8_t)s;

        plan.mode7RGBPartitionEnabled = plan.mode7RGBAPartitionEnabled & ~plan.mode3PartitionEnabled;
        return true;
    }

    void bc7_plan_from_quality(BC7PlanPOD &plan, int quality)
    {
        quality = std::min(100, std::max(1, quality));

        BC7FineTuningPOD ft;
        memset(&ft, 0, sizeof(ft));

        const unsigned short *lists[2] = { kBC7PrioRGB, kBC7PrioRGBA };
        const int counts[2] = { CVTT_BC7_NUM_PRIO_RGB * quality / 100, CVTT_BC7_NUM_PRIO_RGBA * quality / 100 };
        for (int li = 0; li < 2; li++)
            for (int i = 0; i < counts[li]; i++)
            {
                // tools/gen_tables.py: (seedPoints - 1) << 9 | mode << 6 | (partition, or rotation | indexSelector << 2)
                const unsigned code = lists[li][i];
                const uint8_t sp = (uint8_t)(((code >> 9) & 3) + 1);
                const int sub = code & 63;
                switch ((code >> 6) & 7)
                {
                case 0: ft.mode0SP[sub] = sp; break;
                case 1: ft.mode1SP[sub] = sp; break;
                case 2: ft.mode2SP[sub] = sp; break;
                case 3: ft.mode3SP[sub] = sp; break;
                case 4: ft.mode4SP[sub & 3][(sub >> 2) & 1] = sp; break;
                case 5: ft.mode5SP[sub & 3] = sp; break;
                case 6: ft.mode6SP = sp; break;
                case 7: ft.mode7SP[sub] = sp; break;
                }
            }
        bc7_plan_from_fine_tuning(plan, ft);
    }

    // ---------------------------------------------------------------------------------------------------
    // command stream

    namespace
    {
        struct Run { int mode, seeds, slot; };

        // PAIR2: search subset B of the modes with four parity combinations (3, 7) as two tasks of one parity pair each
        // (shorter task phases, the subset's fit is repeated) or as one
        const bool kBC7SplitWideRuns = true;
        // pair form: three-subset modes as SHAPE commands for the anchors + TRIPLE commands (A/B: CVTTB200_BC7_TRIPLE=0 keeps SHAPE / EVAL)
        const bool kBC7TripleCommands = getenv("CVTTB200_BC7_TRIPLE") ? atoi(getenv("CVTTB200_BC7_TRIPLE")) != 0 : true;

        void emit_shape(std::vector<uint32_t> &cmds, const BC7PlanPOD &plan, int shape, const std::vector<Run> &runs)
        {
            bool inRGBList = false, inRGBAList = false, anyRGB = false, anyRGBA = false;
            for (int i = 0; i < plan.rgbNumShapesToEvaluate; i++)
                inRGBList |= (plan.rgbShapeList[i] == shape);
            for (int i = 0; i < plan.rgbaNumShapesToEvaluate; i++)
                inRGBAList |= (plan.rgbaShapeList[i] == shape);
            for (size_t r = 0; r < runs.size(); r++)
                (runs[r].mode < 4 ? anyRGB : anyRGBA) = true;

            // the RGB fit is needed by RGB-mode runs and, through ExpandTo<4>, by RGBA runs of opaque groups
            const bool listedRGB = inRGBList && (anyRGB || anyRGBA);
            const unsigned mask = kBC7ShapeMask[shape];
            cmds.push_back(kCmdShape | ((uint32_t)runs.size() << 8) | ((uint32_t)listedRGB << 16) | ((uint32_t)inRGBAList << 17) | ((uint32_t)anyRGBA << 18));
            cmds.push_back(mask | ((uint32_t)popcount16(mask) << 16));
            for (size_t r = 0; r < runs.size(); r++)
                cmds.push_back((uint32_t)runs[r].mode | ((uint32_t)std::min(runs[r].seeds, 4) << 4) | ((uint32_t)runs[r].slot << 8));
        }

        void emit_eval(std::vector<uint32_t> &cmds, int mode, int partition, int numSubsets, const int *slots)
        {
            cmds.push_back(kCmdEval | ((uint32_t)mode << 8) | ((uint32_t)partition << 16) | ((uint32_t)numSubsets << 24));
            uint32_t w = 0;
            for (int s = 0; s < numSubsets; s++)
                w |= (uint32_t)slots[s] << (8 * s);
            cmds.push_back(w);
        }

        // one PAIR2 command (bc7_core.cuh): the larger subset of the partition is A
        void emit_pair2(std::vector<uint32_t> &cmds, const BC7PlanPOD &plan, int p, bool m1, bool m3, bool m7, bool splitWideRuns)
        {
            const uint8_t *spRGB = plan.seedPointsForShapeRGB, *spRGBA = plan.seedPointsForShapeRGBA;
            const int shapes[2] = { kBC7Shapes2[p * 2], kBC7Shapes2[p * 2 + 1] };
            const unsigned masks[2] = { kBC7ShapeMask[shapes[0]], kBC7ShapeMask[shapes[1]] };
            const int a = popcount16(masks[1]) > popcount16(masks[0]) ? 1 : 0, b = 1 - a;
            bool listedRGB[2] = { false, false }, listedRGBA[2] = { false, false };
            for (int k = 0; k < 2; k++)
            {
                for (int i = 0; i < plan.rgbNumShapesToEvaluate; i++)
                    listedRGB[k] |= (plan.rgbShapeList[i] == shapes[k]);
                for (int i = 0; i < plan.rgbaNumShapesToEvaluate; i++)
                    listedRGBA[k] |= (plan.rgbaShapeList[i] == shapes[k]);
            }
            const int nRuns = (m1 ? 1 : 0) + (m3 ? 1 : 0) + (m7 ? 1 : 0);
            cmds.push_back(kCmdPair2 | ((uint32_t)nRuns << 8) | ((uint32_t)listedRGB[a] << 16) | ((uint32_t)listedRGBA[a] << 17) | ((uint32_t)m7 << 18) |
                           ((uint32_t)listedRGB[b] << 19) | ((uint32_t)listedRGBA[b] << 20) | ((uint32_t)a << 21) | ((uint32_t)(splitWideRuns ? 1 : 0) << 22) | ((uint32_t)p << 24));
            cmds.push_back(masks[a] | ((uint32_t)popcount16(masks[a]) << 16));
            cmds.push_back(masks[b] | ((uint32_t)popcount16(masks[b]) << 16));
            // run word: mode | seeds of subset A << 4 | seeds of subset B << 8
            if (m1) cmds.push_back(1u | ((uint32_t)std::min<int>(spRGB[shapes[a]], 4) << 4) | ((uint32_t)std::min<int>(spRGB[shapes[b]], 4) << 8));
            if (m3) cmds.push_back(3u | ((uint32_t)std::min<int>(spRGB[shapes[a]], 4) << 4) | ((uint32_t)std::min<int>(spRGB[shapes[b]], 4) << 8));
            if (m7) cmds.push_back(7u | ((uint32_t)std::min<int>(spRGBA[shapes[a]], 4) << 4) | ((uint32_t)std::min<int>(spRGBA[shapes[b]], 4) << 8));
        }
    }

    // index (0..2) of the partition's largest subset, the anchor of its TRIPLE command (the first one on ties)
    static int triple_anchor(int p)
    {
        int a = 0;
        for (int k = 1; k < 3; k++)
            if (popcount16(kBC7ShapeMask[kBC7Shapes3[p * 3 + k]]) > popcount16(kBC7ShapeMask[kBC7Shapes3[p * 3 + a]]))
                a = k;
        return a;
    }

    // one TRIPLE command (bc7_core.cuh); slotA0 / slotA2: the result slots of the anchor's mode-0 / mode-2 search
    static void emit_triple(std::vector<uint32_t> &cmds, const BC7PlanPOD &plan, int p, bool m0, bool m2, int slotA0, int slotA2)
    {
        const uint8_t *spRGB = plan.seedPointsForShapeRGB;
        const int anchor = triple_anchor(p);
        int others[2], n = 0;
        for (int k = 0; k < 3; k++)
            if (k != anchor)
                others[n++] = k;
        const int shapeB = kBC7Shapes3[p * 3 + others[0]], shapeC = kBC7Shapes3[p * 3 + others[1]];
        bool listedB = false, listedC = false;
        for (int i = 0; i < plan.rgbNumShapesToEvaluate; i++)
        {
            listedB |= (plan.rgbShapeList[i] == shapeB);
            listedC |= (plan.rgbShapeList[i] == shapeC);
        }
        const int nRuns = (m0 ? 1 : 0) + (m2 ? 1 : 0);
        cmds.push_back(kCmdTriple | ((uint32_t)nRuns << 8) | ((uint32_t)anchor << 16) | ((uint32_t)listedB << 18) | ((uint32_t)listedC << 19) | ((uint32_t)p << 24));
        cmds.push_back(kBC7ShapeMask[shapeB] | ((uint32_t)popcount16(kBC7ShapeMask[shapeB]) << 16) | ((uint32_t)others[0] << 24));
        cmds.push_back(kBC7ShapeMask[shapeC] | ((uint32_t)popcount16(kBC7ShapeMask[shapeC]) << 16) | ((uint32_t)others[1] << 24));
        const uint32_t seedWord = ((uint32_t)std::min<int>(spRGB[shapeB], 4) << 4) | ((uint32_t)std::min<int>(spRGB[shapeC], 4) << 8);
        if (m0) cmds.push_back(0u | seedWord | ((uint32_t)slotA0 << 16));
        if (m2) cmds.push_back(2u | seedWord | ((uint32_t)slotA2 << 16));
    }

    // SHAPE commands for the anchors of the given three-subset partitions (each distinct shape once, slots from nextSlot up),
    // then their TRIPLE commands, adjacent so that the kernel takes them in groups
    static void emit_three_subset_block(std::vector<uint32_t> &cmds, const BC7PlanPOD &plan, const bool enabled[2][64], int &nextSlot)
    {
        const uint8_t *spRGB = plan.seedPointsForShapeRGB;
        int slotOf[2][243];
        for (int i = 0; i < 243; i++)
            slotOf[0][i] = slotOf[1][i] = -1;
        for (int p = 0; p < 64; p++)
            for (int m = 0; m < 2; m++)
                if (enabled[m][p])
                    slotOf[m][kBC7Shapes3[p * 3 + triple_anchor(p)]] = 0;
        for (int shape = 0; shape < 243; shape++)
        {
            std::vector<Run> runs;
            if (slotOf[0][shape] == 0)
                runs.push_back(Run{ 0, spRGB[shape], slotOf[0][shape] = nextSlot++ });
            if (slotOf[1][shape] == 0)
                runs.push_back(Run{ 2, spRGB[shape], slotOf[1][shape] = nextSlot++ });
            if (!runs.empty())
                emit_shape(cmds, plan, shape, runs);
        }
        for (int p = 0; p < 64; p++)
            if (enabled[0][p] || enabled[1][p])
            {
                const int shapeA = kBC7Shapes3[p * 3 + triple_anchor(p)];
                emit_triple(cmds, plan, p, enabled[0][p], enabled[1][p], slotOf[0][shapeA], slotOf[1][shapeA]);
            }
    }

    // kBC7StreamSplit: the stream for SMALL calls.  The search of a block is dealt out to `slices` CTAs (on as many SMs) in
    // independent units -- mode 6, every mode-4/5 rotation, every two-subset partition (one PAIR2 command), every three-subset
    // partition (one TRIPLE command; the anchors of a slice's partitions are searched once per slice, in front of its TRIPLE
    // commands, so no unit reads another slice's results).  Units go longest first to the least loaded slice.  Layout: cmds[s] =
    // offset of slice s's sub-stream (each ends with END), s < slices.  The launch reduces the slices' winners by (error,
    // reference key): bc7_candidate_merge.
    static int compile_split(const BC7PlanPOD &plan, std::vector<uint32_t> &cmds, int slices)
    {
        struct Unit { int cost, kind, partition; bool m0, m2; std::vector<uint32_t> words; };      // kind: 0 mode 6, 1 modes 4 / 5, 2 three subsets, 3 two subsets
        std::vector<Unit> units;
        const uint8_t *spRGB = plan.seedPointsForShapeRGB, *spRGBA = plan.seedPointsForShapeRGBA;
        // trial passes of one shape: (parity combinations, two per pass) x seeds
        auto passes = [](int mode, int seeds) { seeds = std::min(seeds, 4); return mode == 2 ? (seeds + 1) / 2 : (mode == 1 ? seeds : 2 * seeds); };

        if (plan.mode6Enabled && spRGBA[0])
        {
            Unit u;
            std::vector<Run> runs(1, Run{ 6, spRGBA[0], 0 });
            emit_shape(u.words, plan, 0, runs);
            const int slots[1] = { 0 };
            emit_eval(u.words, 6, 0, 1, slots);
            u.cost = 20 * passes(6, spRGBA[0]) + 16;
            u.kind = 0;
            units.push_back(u);
        }
        for (int mode = 4; mode <= 5; mode++)
            for (int rotation = 0; rotation < 4; rotation++)
                for (int isel = 0; isel < (mode == 4 ? 2 : 1); isel++)
                {
                    const int seeds = std::min<int>(4, mode == 4 ? plan.mode4SP[rotation][isel] : plan.mode5SP[rotation]);
                    if (seeds <= 0)
                        continue;
                    Unit u;
                    u.words.push_back(kCmdDual | ((uint32_t)mode << 8) | ((uint32_t)rotation << 16) | ((uint32_t)isel << 20) | ((uint32_t)seeds << 24));
                    u.cost = 16 * 2 * seeds + 16;
                    u.kind = 1;
                    units.push_back(u);
                }
        for (int p = 0; p < 64; p++)
        {
            const int shapes[2] = { kBC7Shapes2[p * 2], kBC7Shapes2[p * 2 + 1] };
            const bool rgbSearched = spRGB[shapes[0]] && spRGB[shapes[1]];
            const bool m1 = ((plan.mode1PartitionEnabled >> p) & 1) && rgbSearched;
            const bool m3 = ((plan.mode3PartitionEnabled >> p) & 1) && rgbSearched;
            const bool m7 = spRGBA[shapes[0]] && spRGBA[shapes[1]];
            if (!m1 && !m3 && !m7)
                continue;
            Unit u;
            emit_pair2(u.words, plan, p, m1, m3, m7, kBC7SplitWideRuns);
            u.cost = 16;
            u.kind = 3;
            for (int k = 0; k < 2; k++)
                u.cost += popcount16(kBC7ShapeMask[shapes[k]]) * ((m1 ? passes(1, spRGB[shapes[k]]) : 0) + (m3 ? passes(3, spRGB[shapes[k]]) : 0) + (m7 ? passes(7, spRGBA[shapes[k]]) : 0));
            units.push_back(u);
        }
        for (int p = 0; p < 64; p++)
        {
            const bool searched = spRGB[kBC7Shapes3[p * 3]] && spRGB[kBC7Shapes3[p * 3 + 1]] && spRGB[kBC7Shapes3[p * 3 + 2]];
            const bool m0 = p < 16 && ((plan.mode0PartitionEnabled >> p) & 1) && searched;
            const bool m2 = ((plan.mode2PartitionEnabled >> p) & 1) && searched;
            if (!m0 && !m2)
                continue;
            // its words are made per slice (below): the anchors of a slice's partitions are searched once, their TRIPLE commands follow
            Unit u;
            u.kind = 2;
            u.partition = p;
            u.m0 = m0;
            u.m2 = m2;
            const int shapeA = kBC7Shapes3[p * 3 + triple_anchor(p)];
            u.cost = 24 + popcount16(kBC7ShapeMask[shapeA]) * ((m0 ? passes(0, spRGB[shapeA]) : 0) + (m2 ? passes(2, spRGB[shapeA]) : 0)) * 5 / 4;
            units.push_back(u);
        }

        // Longest unit first onto the least loaded slice.  Within a slice the cheap one-subset and three-subset units go first:
        // they give the block an error to beat, which lets the PAIR2 commands behind them drop second subsets.  (Starting
        // every slice with a mode-6 search of its own for the same reason was measured: no gain at 3 slices, 15 % slower at 48.)
        std::vector<size_t> order(units.size());
        for (size_t i = 0; i < order.size(); i++)
            order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return units[a].cost > units[b].cost; });
        std::vector<int> load((size_t)slices, 0);
        std::vector<std::vector<size_t> > assigned((size_t)slices);
        for (size_t k = 0; k < order.size(); k++)
        {
            const Unit &u = units[order[k]];
            const size_t slice = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
            load[slice] += u.cost;
            assigned[slice].push_back(order[k]);
        }
        std::vector<std::vector<uint32_t> > streams((size_t)slices);
        int maxSlot = 12;
        for (size_t s = 0; s < (size_t)slices; s++)
        {
            std::stable_sort(assigned[s].begin(), assigned[s].end(), [&](size_t a, size_t b) { return units[a].kind < units[b].kind; });
            bool enabled[2][64];
            memset(enabled, 0, sizeof(enabled));
            bool anyThree = false, threeDone = false;
            for (size_t k = 0; k < assigned[s].size(); k++)
                if (units[assigned[s][k]].kind == 2)
                {
                    enabled[0][units[assigned[s][k]].partition] = units[assigned[s][k]].m0;
                    enabled[1][units[assigned[s][k]].partition] = units[assigned[s][k]].m2;
                    anyThree = true;
                }
            for (size_t k = 0; k < assigned[s].size(); k++)
            {
                const Unit &u = units[assigned[s][k]];
                if (u.kind == 2)
                {
                    if (anyThree && !threeDone)
                    {
                        int nextSlot = 6;
                        emit_three_subset_block(streams[s], plan, enabled, nextSlot);
                        maxSlot = std::max(maxSlot, nextSlot);
                        threeDone = true;
                    }
                    continue;
                }
                streams[s].insert(streams[s].end(), u.words.begin(), u.words.end());
            }
        }
        cmds.assign((size_t)slices, 0u);
        for (size_t s = 0; s < (size_t)slices; s++)
        {
            cmds[s] = (uint32_t)cmds.size();
            cmds.insert(cmds.end(), streams[s].begin(), streams[s].end());
            cmds.push_back(kCmdEnd);
        }
        return maxSlot;
    }

    int bc7_compile_plan(const BC7PlanPOD &plan, std::vector<uint32_t> &cmds, int form, int slices)
    {
        if (form == kBC7StreamSplit)
            return compile_split(plan, cmds, std::max(1, slices));
        const bool pairCommands = (form == kBC7StreamPair);
        cmds.clear();
        const uint8_t *spRGB = plan.seedPointsForShapeRGB, *spRGBA = plan.seedPointsForShapeRGBA;
        int maxSlot = 0;

        // The one-subset candidates go first (mode 6, then modes 4 and 5), then the three-subset modes: they are the cheap part
        // of the search and give every block a good error to beat, which is what lets the PAIR2 commands of the two-subset
        // modes drop most second subsets.  The order of commands is free: the winner is the lexicographic (error, reference
        // sequence) minimum.  Slots 0-5 are recycled (mode 6, then the two-subset partitions), the rest belong to modes 0 / 2.

        // Mode 6: the whole block (shape 0)
        if (plan.mode6Enabled && spRGBA[0])
        {
            std::vector<Run> runs(1, Run{ 6, spRGBA[0], 0 });
            emit_shape(cmds, plan, 0, runs);
            const int slots[1] = { 0 };
            emit_eval(cmds, 6, 0, 1, slots);
            maxSlot = std::max(maxSlot, 1);
        }

        // Modes 4 and 5 (BC67.cpp:1675-1683, 1735-1742)
        for (int mode = 4; mode <= 5; mode++)
            for (int rotation = 0; rotation < 4; rotation++)
                for (int isel = 0; isel < (mode == 4 ? 2 : 1); isel++)
                {
                    const int seeds = std::min<int>(4, mode == 4 ? plan.mode4SP[rotation][isel] : plan.mode5SP[rotation]);
                    if (seeds > 0)
                        cmds.push_back(kCmdDual | ((uint32_t)mode << 8) | ((uint32_t)rotation << 16) | ((uint32_t)isel << 20) | ((uint32_t)seeds << 24));
                }

        // Three-subset modes 0 and 2: shapes are shared between partitions, so each needed shape is searched once
        // and kept in its own slot until the partitions are scanned.
        {
            int slotOf[2][243];
            for (int i = 0; i < 243; i++)
                slotOf[0][i] = slotOf[1][i] = -1;
            bool enabled[2][64];
            for (int p = 0; p < 64; p++)
            {
                const bool searched = spRGB[kBC7Shapes3[p * 3]] && spRGB[kBC7Shapes3[p * 3 + 1]] && spRGB[kBC7Shapes3[p * 3 + 2]];
                enabled[0][p] = p < 16 && ((plan.mode0PartitionEnabled >> p) & 1) && searched;
                enabled[1][p] = ((plan.mode2PartitionEnabled >> p) & 1) && searched;
            }
            int nextSlot = 6;
            if (pairCommands && kBC7TripleCommands)
            {
                // TRIPLE commands (bc7_core.cuh): only each partition's largest subset, the anchor, is searched for every block
                emit_three_subset_block(cmds, plan, enabled, nextSlot);
                maxSlot = std::max(maxSlot, nextSlot);
            }
            else
            {
            for (int m = 0; m < 2; m++)
                for (int p = 0; p < 64; p++)
                    if (enabled[m][p])
                        for (int k = 0; k < 3; k++)
                            slotOf[m][kBC7Shapes3[p * 3 + k]] = 0;
            for (int shape = 0; shape < 243; shape++)
            {
                std::vector<Run> runs;
                if (slotOf[0][shape] == 0)
                    runs.push_back(Run{ 0, spRGB[shape], slotOf[0][shape] = nextSlot++ });
                if (slotOf[1][shape] == 0)
                    runs.push_back(Run{ 2, spRGB[shape], slotOf[1][shape] = nextSlot++ });
                if (!runs.empty())
                    emit_shape(cmds, plan, shape, runs);
            }
            for (int m = 0; m < 2; m++)
                for (int p = 0; p < 64; p++)
                    if (enabled[m][p])
                    {
                        const int slots[3] = { slotOf[m][kBC7Shapes3[p * 3]], slotOf[m][kBC7Shapes3[p * 3 + 1]], slotOf[m][kBC7Shapes3[p * 3 + 2]] };
                        emit_eval(cmds, m ? 2 : 0, p, 3, slots);
                    }
            maxSlot = std::max(maxSlot, nextSlot);
            }
        }

        // Two-subset modes 1, 3, 7, partition-major: every two-subset shape belongs to exactly one partition.  A (mode,
        // partition) whose subsets are not all searched can never win in the reference (its total is >= FLT_MAX), so it is not
        // emitted at all.  Mode 7 scans all 64 partitions whatever mode7RGBAPartitionEnabled says (the reference assigns a
        // misspelt variable, BC67.cpp:1593-1596).
        //   pairCommands: one PAIR2 command per partition -- the larger subset (A) is searched for every block, the other one
        //   (B) only for the blocks whose err(A) leaves room below their best (bc7_core.cuh);
        //   otherwise SHAPE, SHAPE, EVAL... through six recycled slots (the kernels with per-trial group votes and the
        //   single-colour candidates need every block's own thread to walk both subsets).
        for (int p = 0; p < 64; p++)
        {
            const int shapes[2] = { kBC7Shapes2[p * 2], kBC7Shapes2[p * 2 + 1] };
            const bool rgbSearched = spRGB[shapes[0]] && spRGB[shapes[1]];
            const bool m1 = ((plan.mode1PartitionEnabled >> p) & 1) && rgbSearched;
            const bool m3 = ((plan.mode3PartitionEnabled >> p) & 1) && rgbSearched;
            const bool m7 = spRGBA[shapes[0]] && spRGBA[shapes[1]];
            if (!m1 && !m3 && !m7)
                continue;
            if (pairCommands)
            {
                emit_pair2(cmds, plan, p, m1, m3, m7, kBC7SplitWideRuns);
                continue;
            }
            for (int s = 0; s < 2; s++)
            {
                std::vector<Run> runs;
                if (m1) runs.push_back(Run{ 1, spRGB[shapes[s]], 0 + s });
                if (m3) runs.push_back(Run{ 3, spRGB[shapes[s]], 2 + s });
                if (m7) runs.push_back(Run{ 7, spRGBA[shapes[s]], 4 + s });
                emit_shape(cmds, plan, shapes[s], runs);
            }
            const int s1[2] = { 0, 1 }, s3[2] = { 2, 3 }, s7[2] = { 4, 5 };
            if (m1) emit_eval(cmds, 1, p, 2, s1);
            if (m3) emit_eval(cmds, 3, p, 2, s3);
            if (m7) emit_eval(cmds, 7, p, 2, s7);
            maxSlot = std::max(maxSlot, 6);
        }

        cmds.push_back(kCmdEnd);
        return maxSlot;
    }

    // ---------------------------------------------------------------------------------------------------
    // constants

    static QuantConst make_quant(int bits, bool withP, int unqBits)
    {
        QuantConst q;
        memset(&q, 0, sizeof(q));
        if (withP)
        {
            // QuantizeP: N = c * (2^(bits+1) - 1) + addend(p); v = N >> 9; out = 2v + p
            const int K = (1 << (bits + 1)) - 1;
            const int addend[2] = { 255, (1 << (8 - bits)) - 1 };
            q.qMul = (float)K / 512.0f;
            for (int p = 0; p < 2; p++)
                q.qAdd[p] = (float)(2 * addend[p] - 511) / 1024.0f;
            q.pMul = 2.0f;
            q.pAdd = 1.0f;
        }
        else
        {
            // Quantize: N = c * (2^bits - 1) + 127 + 2^(7-bits); v = N >> 8
            const int K = (1 << bits) - 1;
            const int addend = 127 + (1 << (7 - bits));
            q.qMul = (float)K / 256.0f;
            q.qAdd[0] = q.qAdd[1] = (float)(2 * addend - 255) / 512.0f;
            q.pMul = 1.0f;
            q.pAdd = 0.0f;
        }
        if (unqBits)
        {
            // Unquantize: out = v * 2^(8-b) + floor(v / 2^(2b-8))
            const int s = 2 * unqBits - 8;
            q.hasUnq = 1;
            q.uMul = (float)(1 << (8 - unqBits));
            q.uScale = 1.0f / (float)(1 << s);
            q.uOff = -(float)((1 << s) - 1) / (float)(1 << (s + 1));
        }
        return q;
    }

    void bc7_fill_params(BC7Params &P, const OptionsPOD &options, const BC7PlanPOD &plan, const float rcpN[17])
    {
        memset(&P, 0, sizeof(P));

        // Util::FillWeights, ConvectionKernels_Util.cpp:62-73
        if (options.flags & kFlag_Uniform)
            P.w[0] = P.w[1] = P.w[2] = P.w[3] = 1.0f;
        else
        {
            P.w[0] = options.redWeight;
            P.w[1] = options.greenWeight;
            P.w[2] = options.blueWeight;
            P.w[3] = options.alphaWeight;
        }
        for (int ch = 0; ch < 4; ch++)
        {
            P.wSq[ch] = P.w[ch] * P.w[ch];
            P.rcpW[ch] = (P.w[ch] != 0.0f) ? 1.0f / P.w[ch] : 1.0f;    // EndpointRefiner.h:52-57
        }
        for (int n = 0; n < 17; n++)
            P.rcpN[n] = rcpN[n];

        for (int bits = 2; bits <= 4; bits++)
        {
            IndexConst &ic = P.ic[bits - 2];
            const int range = 1 << bits;
            ic.maxValue = (float)(range - 1);
            ic.wScale = 64.0f / (float)(range - 1);
            ic.rcpMaxIndex = 1.0f / (float)(range - 1);
            for (int tweak = 0; tweak < 4; tweak++)
            {
                // Util::ComputeTweakFactors, ConvectionKernels_Util.cpp:75-85
                const int totalUnits = range - 1;
                const int minOutsideUnits = (tweak >> 1) & 1, maxOutsideUnits = tweak & 1;
                const int insideUnits = totalUnits - minOutsideUnits - maxOutsideUnits;
                ic.tweak[tweak][0] = -(float)minOutsideUnits / (float)insideUnits;
                ic.tweak[tweak][1] = (float)maxOutsideUnits / (float)insideUnits + 1.0f;
            }
        }

        // g_modes (BC67.cpp:108-119) and CompressEndpoints0-7 (BC67.cpp:862-938)
        struct { int bits, withP, unq, parityMax, sharedP, indexBits; } const modes[8] =
        {
            { 4, 1, 5, 4, 0, 3 },
            { 6, 1, 7, 2, 1, 3 },
            { 5, 0, 5, 1, 0, 2 },
            { 7, 1, 0, 4, 0, 2 },
            { 5, 0, 5, 1, 0, 2 },   // mode 4 RGB; index precision is chosen per index selector
            { 7, 0, 7, 1, 0, 2 },   // mode 5 RGB
            { 7, 1, 0, 4, 0, 4 },
            { 5, 1, 6, 4, 0, 2 },
        };
        for (int m = 0; m < 8; m++)
        {
            P.mc[m].q = make_quant(modes[m].bits, modes[m].withP != 0, modes[m].unq);
            P.mc[m].parityBitMax = modes[m].parityMax;
            P.mc[m].sharedP = modes[m].sharedP;
            P.mc[m].indexBits = modes[m].indexBits;
        }
        P.alphaQ4 = make_quant(6, false, 6);

        P.mode7RGBPartitionEnabled = plan.mode7RGBPartitionEnabled;
        P.flags = options.flags;
        P.refineRounds = std::max(1, options.refineRoundsBC7);     // BC67.cpp:1046-1047
    }

    const BC7PackTables &bc7_pack_tables()
    {
        static BC7PackTables T;
        static bool ready = false;
        if (!ready)
        {
            for (int i = 0; i < 64; i++)
            {
                T.partitionMask2[i] = kBC7PartitionMask2[i];
                T.partitionMap3[i] = kBC7PartitionMap3[i];
                T.fixup2[i] = kBC7Fixup2[i];
            }
            for (int i = 0; i < 128; i++)
                T.fixup3[i] = kBC7Fixup3[i];
            ready = true;
        }
        return T;
    }
}
