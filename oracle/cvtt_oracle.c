/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's BC7 encode path.
 *
 * Plain scalar C, one block at a time, written to follow the reference (elasota/ConvectionKernels,
 * /root/reference) statement by statement so that its output is bit-identical to the reference's SSE2
 * ParallelMath back end.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this file's library; the product path (convectionkernels_b200/csrc) never does.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this restatement is
 * pinned against the reference itself, compiled unmodified into oracle/_ref/libcvtt_ref.so by
 * oracle/Makefile (tests/test_oracle.py compares the two block by block; tests/golden/ holds fixtures
 * generated from oracle/_ref by tools/make_golden.py).
 *
 * Numerical contract reproduced here (reference ConvectionKernels_ParallelMath.h:64-1279, SSE2 half):
 *   - fp32 arithmetic, one rounding per operation (no FMA contraction: build with -ffp-contract=off);
 *   - Min/Max/Clamp have _mm_min_ps/_mm_max_ps operand semantics (second operand wins ties) :522-559;
 *   - RoundAndConvertToU15 = cvtps2dq under round-to-nearest-even + signed saturating pack :935-945;
 *   - Reciprocal = the host's _mm_rcp_ps :569-575 (so the result depends on the host CPU model);
 *   - UInt15/UInt16 lanes wrap at 16 bits;
 *   - automatic variables the reference leaves uninitialised are ZERO here, matching the oracle build's
 *     -ftrivial-auto-var-init=zero (matters for BC7 mode 6 on opaque groups, see try_single_plane).
 *
 * The reference processes 8 blocks per call in SIMD lanes; two decisions are taken across the 8 lanes
 * (ConvectionKernels_BC67.cpp:1069,1072) and change per-block results, so blocks are encoded here in
 * groups of 8 with those two flags computed per group.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <xmmintrin.h>
#include <emmintrin.h>

#include "bc7_tables.inc"

/* ---- public structs: layouts of cvtt::Options / cvtt::BC7EncodingPlan (ConvectionKernels.h:73-199) ---- */
typedef struct
{
    uint32_t flags;
    float threshold, redWeight, greenWeight, blueWeight, alphaWeight;
    int refineRoundsBC7, refineRoundsBC6H, refineRoundsIIC, refineRoundsS3TC, seedPoints;
} oracle_options;

typedef struct
{
    uint64_t mode1PartitionEnabled, mode2PartitionEnabled, mode3PartitionEnabled;
    uint16_t mode0PartitionEnabled;
    uint64_t mode7RGBAPartitionEnabled, mode7RGBPartitionEnabled;
    uint8_t mode4SP[4][2];
    uint8_t mode5SP[4];
    uint8_t mode6Enabled; /* bool */
    uint8_t seedPointsForShapeRGB[243];
    uint8_t seedPointsForShapeRGBA[129];
    uint8_t rgbaShapeList[129];
    uint8_t rgbaNumShapesToEvaluate;
    uint8_t rgbShapeList[243];
    uint8_t rgbNumShapesToEvaluate;
} oracle_bc7_plan;

enum
{
    FLAG_BC7_FastIndexing = 0x008,
    FLAG_BC7_TrySingleColor = 0x010,
    FLAG_BC7_RespectPunchThrough = 0x020,
    FLAG_Uniform = 0x200
};

#define MAX_TWEAK_ROUNDS 4 /* ConvectionKernels_BC67.h:41 */

/* ---- ParallelMath lane semantics ---- */
static float pm_min(float a, float b) { return (a < b) ? a : b; }   /* _mm_min_ps(a, b) */
static float pm_max(float a, float b) { return (a > b) ? a : b; }   /* _mm_max_ps(a, b) */
static float pm_clamp(float v, float lo, float hi) { return pm_max(pm_min(v, hi), lo); } /* ParallelMath.h:561-567 */
/* _mm_rcp_ps differs between CPU vendors/models.  On the BC7 path its argument is always a pixel count 1..16, so a
 * 17-entry table recorded with a golden fixture can stand in for the instruction (tests/golden). */
static float g_rcp_override[17];
static int g_rcp_overridden = 0;
static float pm_rcp(float v)                                                           /* :569-575 */
{
    if (g_rcp_overridden && v >= 1.0f && v <= 16.0f && v == (float)(int)v)
        return g_rcp_override[(int)v];
    return _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps(v)));
}
void cvtt_oracle_set_rcp_table(const float *t)
{
    g_rcp_overridden = (t != NULL);
    if (t)
        memcpy(g_rcp_override, t, sizeof(g_rcp_override));
}
static void pm_safe_denominator(float *v) { if (*v == 0.0f) *v = 1.0f; }               /* :472-475 */

static uint16_t pm_round_u15(float v)                                                  /* :935-945 */
{
    int i = _mm_cvtss_si32(_mm_set_ss(v)); /* MXCSR rounding = nearest even (RoundTowardNearestForScope) */
    if (i > 32767) i = 32767;
    if (i < -32768) i = -32768;
    return (uint16_t)(int16_t)i;
}

/* ---- shapes: a shape is a 16-bit pixel mask, its pixel order is ascending bit order ---- */
static int shape_pixels(int shape, uint8_t px[16])
{
    unsigned m = kBC7ShapeMask[shape];
    int n = 0;
    for (int p = 0; p < 16; p++)
        if (m & (1u << p))
            px[n++] = (uint8_t)p;
    return n;
}

static int g_shape_start[CVTT_BC7_NUM_SHAPES]; /* offset of the shape in the reference's g_fragments array */
static int g_tables_ready = 0;

static void init_tables(void)
{
    if (g_tables_ready)
        return;
    int off = 0;
    for (int s = 0; s < CVTT_BC7_NUM_SHAPES; s++)
    {
        g_shape_start[s] = off;
        off += __builtin_popcount(kBC7ShapeMask[s]);
    }
    g_tables_ready = 1;
}

/* ---- Util::ComputeTweakFactors, ConvectionKernels_Util.cpp:75-85 ---- */
static void compute_tweak_factors(int tweak, int range, float *outFactors)
{
    int totalUnits = range - 1;
    int minOutsideUnits = ((tweak >> 1) & 1);
    int maxOutsideUnits = (tweak & 1);
    int insideUnits = totalUnits - minOutsideUnits - maxOutsideUnits;

    outFactors[0] = -(float)minOutsideUnits / (float)insideUnits;
    outFactors[1] = (float)maxOutsideUnits / (float)insideUnits + 1.0f;
}

/* ---- UnfinishedEndpoints<N>, ConvectionKernels_UnfinishedEndpoints.h ---- */
typedef struct { float base[4], offset[4]; } ufep_t;

static void ufep_finish_ldr(const ufep_t *u, int nch, int tweak, int range, uint16_t *ep0, uint16_t *ep1) /* :75-91 */
{
    float tf[2];
    compute_tweak_factors(tweak, range, tf);
    for (int ch = 0; ch < nch; ch++)
    {
        float ep0f = pm_clamp(u->base[ch] + u->offset[ch] * tf[0], 0.0f, 255.0f);
        float ep1f = pm_clamp(u->base[ch] + u->offset[ch] * tf[1], 0.0f, 255.0f);
        ep0[ch] = pm_round_u15(ep0f);
        ep1[ch] = pm_round_u15(ep1f);
    }
}

/* ---- EndpointSelector<N, 8> + PackedCovarianceMatrix<N>, ConvectionKernels_EndpointSelector.h,
 *      ConvectionKernels_PackedCovarianceMatrix.h.  weight is always 1.0f on the BC7 path. ---- */
static void endpoint_selector(const float pw[16][4], int chBase, const uint8_t *pxlist, int n, int nch, const float *channelWeights, ufep_t *out)
{
    float centroid[4] = { 0, 0, 0, 0 }, direction[4] = { 0, 0, 0, 0 };
    float cov[10];
    float weightTotal = 0.0f, minDist = FLT_MAX, maxDist = -FLT_MAX;
    const float weight = 1.0f;
    const int pyramid = nch * (nch + 1) / 2;
    (void)chBase;

    for (int i = 0; i < pyramid; i++)
        cov[i] = 0.0f;

    /* pass 0: centroid (:73-86) */
    for (int i = 0; i < n; i++)
    {
        const float *value = pw[pxlist[i]];
        for (int ch = 0; ch < nch; ch++)
            centroid[ch] = centroid[ch] + value[ch] * weight;
        weightTotal = weightTotal + weight;
    }
    {
        float denom = weightTotal;
        pm_safe_denominator(&denom);
        for (int ch = 0; ch < nch; ch++)
            centroid[ch] = centroid[ch] / denom;
    }

    /* pass 1: covariance (:88-95, PackedCovarianceMatrix.h:29-40) */
    for (int i = 0; i < n; i++)
    {
        const float *value = pw[pxlist[i]];
        float diff[4];
        for (int ch = 0; ch < nch; ch++)
            diff[ch] = value[ch] - centroid[ch];
        int index = 0;
        for (int row = 0; row < nch; row++)
            for (int col = 0; col <= row; col++)
            {
                cov[index] = cov[index] + diff[row] * diff[col] * weight;
                index++;
            }
    }

    /* power iteration (:97-130, PackedCovarianceMatrix.h:42-60) */
    {
        float approx[4];
        for (int ch = 0; ch < nch; ch++)
            approx[ch] = 1.0f;

        for (int it = 0; it < 8; it++)
        {
            float product[4];
            for (int row = 0; row < nch; row++)
            {
                float sum = 0.0f;
                int index = (row * (row + 1)) >> 1;
                for (int col = 0; col < nch; col++)
                {
                    sum = sum + approx[col] * cov[index];
                    if (col >= row)
                        index += col + 1;
                    else
                        index++;
                }
                product[row] = sum;
            }

            float largestComponent = product[0];
            for (int ch = 1; ch < nch; ch++)
                largestComponent = pm_max(largestComponent, product[ch]);

            pm_safe_denominator(&largestComponent);
            for (int ch = 0; ch < nch; ch++)
                approx[ch] = product[ch] / largestComponent;
        }

        float approxLen = 0.0f;
        for (int ch = 0; ch < nch; ch++)
            approxLen = approxLen + approx[ch] * approx[ch];
        approxLen = _mm_cvtss_f32(_mm_sqrt_ss(_mm_set_ss(approxLen)));
        pm_safe_denominator(&approxLen);
        for (int ch = 0; ch < nch; ch++)
            direction[ch] = approx[ch] / approxLen;
    }

    /* pass 2: extent along the axis (:132-140) */
    for (int i = 0; i < n; i++)
    {
        const float *value = pw[pxlist[i]];
        float dist = 0.0f;
        for (int ch = 0; ch < nch; ch++)
            dist = dist + direction[ch] * (value[ch] - centroid[ch]);
        minDist = pm_min(minDist, dist);
        maxDist = pm_max(maxDist, dist);
    }

    /* GetEndpoints (:51-70): divides by the raw channel weight */
    for (int ch = 0; ch < nch; ch++)
    {
        float mn = centroid[ch] + direction[ch] * minDist;
        float mx = centroid[ch] + direction[ch] * maxDist;
        out->base[ch] = mn / channelWeights[ch];
        out->offset[ch] = (mx - mn) / channelWeights[ch];
    }
}

/* ---- IndexSelector<N>, ConvectionKernels_IndexSelector.h ---- */
static const uint16_t kWeightReciprocals[17] = /* ConvectionKernels_IndexSelector.cpp:43-62 */
{ 0, 0, 32768, 16384, 10923, 8192, 6554, 5461, 4681, 4096, 3641, 3277, 2979, 2731, 2521, 2341, 2185 };

typedef struct
{
    uint16_t endPoint[2][4];
    float origin[4], axis[4];
    int range;
    float maxValue;
} index_selector;

static void isel_init(index_selector *s, int nch, const float *channelWeights, uint16_t ep[2][4], int range) /* :27-78 */
{
    float epDiffWeighted[4];
    for (int e = 0; e < 2; e++)
        for (int ch = 0; ch < nch; ch++)
            s->endPoint[e][ch] = ep[e][ch];
    s->range = range;
    s->maxValue = (float)(range - 1);

    for (int ch = 0; ch < nch; ch++)
    {
        s->origin[ch] = (float)ep[0][ch];
        float opposingOriginCh = (float)ep[1][ch];
        epDiffWeighted[ch] = (opposingOriginCh - s->origin[ch]) * channelWeights[ch];
    }

    float lenSquared = epDiffWeighted[0] * epDiffWeighted[0];
    for (int ch = 1; ch < nch; ch++)
        lenSquared = lenSquared + epDiffWeighted[ch] * epDiffWeighted[ch];
    pm_safe_denominator(&lenSquared);

    float maxValueDividedByLengthSquared = s->maxValue / lenSquared;
    for (int ch = 0; ch < nch; ch++)
        s->axis[ch] = epDiffWeighted[ch] * channelWeights[ch] * maxValueDividedByLengthSquared;
}

static uint16_t isel_select_index_ldr(const index_selector *s, int nch, const float *pixel) /* :124-131 */
{
    float dist = (pixel[0] - s->origin[0]) * s->axis[0];
    for (int ch = 1; ch < nch; ch++)
        dist = dist + (pixel[ch] - s->origin[ch]) * s->axis[ch];
    return pm_round_u15(pm_clamp(dist, 0.0f, s->maxValue));
}

static void isel_reconstruct_bc7(const index_selector *s, uint16_t index, uint16_t *pixel, int numRealChannels) /* :90-100 */
{
    uint16_t weight = (uint16_t)((uint16_t)((uint16_t)(kWeightReciprocals[s->range] * index) + 256) >> 9);
    for (int ch = 0; ch < numRealChannels; ch++)
    {
        uint16_t ep0f = (uint16_t)((uint16_t)(64 - weight) * s->endPoint[0][ch]);
        uint16_t ep1f = (uint16_t)(weight * s->endPoint[1][ch]);
        pixel[ch] = (uint16_t)((uint16_t)(ep0f + ep1f + 32) >> 6);
    }
}

/* ---- AggregatedError<N> / BCCommon::ComputeErrorLDR, ConvectionKernels_AggregatedError.h,
 *      ConvectionKernels_BCCommon.h:24-43, ParallelMath.h:987-994 ---- */
typedef struct { uint32_t err[4]; } agg_error;

static void agg_init(agg_error *a) { a->err[0] = a->err[1] = a->err[2] = a->err[3] = 0; }

static void compute_error_ldr(const uint16_t *reconstructed, const uint16_t *original, int numRealChannels, agg_error *agg)
{
    for (int ch = 0; ch < numRealChannels; ch++)
    {
        uint16_t diff = (uint16_t)(reconstructed[ch] - original[ch]);
        uint16_t sq = (uint16_t)((uint32_t)diff * (uint32_t)diff); /* _mm_mullo_epi16 */
        agg->err[ch] += sq;
    }
}

static float agg_finalize(const agg_error *a, int nch, uint32_t flags, const float *channelWeightsSq) /* AggregatedError.h:25-46 */
{
    if (flags & FLAG_Uniform)
    {
        uint32_t total = a->err[0];
        for (int ch = 1; ch < nch; ch++)
            total = total + a->err[ch];
        return (float)(int32_t)total;
    }
    else
    {
        float total = (float)(int32_t)a->err[0] * channelWeightsSq[0];
        for (int ch = 1; ch < nch; ch++)
            total = total + (float)(int32_t)a->err[ch] * channelWeightsSq[ch];
        return total;
    }
}

static float compute_error_ldr_simple(uint32_t flags, const uint16_t *reconstructed, const uint16_t *original, int nch, int numRealChannels, const float *channelWeightsSq)
{
    agg_error agg;
    agg_init(&agg);
    compute_error_ldr(reconstructed, original, numRealChannels, &agg);
    return agg_finalize(&agg, nch, flags, channelWeightsSq);
}

/* ---- EndpointRefiner<N>, ConvectionKernels_EndpointRefiner.h ---- */
typedef struct
{
    float tv[4], v[4], tt, t, w;
    int wu;
    float rcpMaxIndex;
    float rcpChannelWeights[4];
} ep_refiner;

static void refiner_init(ep_refiner *r, int nch, int indexRange, const float *channelWeights) /* :38-60 */
{
    for (int ch = 0; ch < nch; ch++)
        r->tv[ch] = r->v[ch] = 0.0f;
    r->tt = r->t = r->w = 0.0f;
    r->rcpMaxIndex = 1.0f / (float)(indexRange - 1);
    for (int ch = 0; ch < nch; ch++)
    {
        r->rcpChannelWeights[ch] = 1.0f;
        if (channelWeights[ch] != 0.0f)
            r->rcpChannelWeights[ch] = 1.0f / channelWeights[ch];
    }
    r->wu = 0;
}

static void refiner_contribute_unweighted_pw(ep_refiner *r, const float *pwFloatPixel, uint16_t index, int numRealChannels) /* :78-92 */
{
    float t = (float)index * r->rcpMaxIndex;
    for (int ch = 0; ch < numRealChannels; ch++)
    {
        float v = pwFloatPixel[ch];
        r->tv[ch] = r->tv[ch] + t * v;
        r->v[ch] = r->v[ch] + v;
    }
    r->tt = r->tt + t * t;
    r->t = r->t + t;
    r->wu++;
}

static void refiner_get_refined_endpoints_ldr(const ep_refiner *r, int nch, uint16_t endPoint[2][4]) /* :99-152 */
{
    float w = r->w + (float)r->wu;
    pm_safe_denominator(&w);
    float wRcp = pm_rcp(w);

    float adenom = (r->tt * w - r->t * r->t) * wRcp;
    int adenomZero = (adenom == 0.0f);
    if (adenomZero)
        adenom = 1.0f;

    for (int ch = 0; ch < nch; ch++)
    {
        float a = (r->tv[ch] - r->t * r->v[ch] * wRcp) / adenom;
        float b = (r->v[ch] - a * r->t) * wRcp;

        float p1 = b;
        float p2 = a + b;

        if (adenomZero)
        {
            p1 = r->v[ch] * wRcp;
            p2 = p1;
        }

        float inverseWeight = r->rcpChannelWeights[ch];
        float e0 = p1 * inverseWeight;
        float e1 = p2 * inverseWeight;
        endPoint[0][ch] = pm_round_u15(pm_clamp(e0, 0.0f, 255.0f));
        endPoint[1][ch] = pm_round_u15(pm_clamp(e1, 0.0f, 255.0f));
    }
}

/* ---- BC7Computer endpoint quantisers, ConvectionKernels_BC67.cpp:829-938 ---- */
static void quantize(uint16_t *color, int bits, int channels) /* :829-833 */
{
    for (int ch = 0; ch < channels; ch++)
        color[ch] = (uint16_t)((uint16_t)((uint16_t)((uint16_t)(color[ch] << bits) - color[ch]) + (uint16_t)(127 + (1 << (7 - bits)))) >> 8);
}

static void quantize_p(uint16_t *color, int bits, uint16_t p, int channels) /* :835-851 */
{
    int16_t addend = p ? (int16_t)((1 << (8 - bits)) - 1) : 255;
    for (int ch = 0; ch < channels; ch++)
    {
        uint16_t ch16 = color[ch];
        ch16 = (uint16_t)((uint16_t)((uint16_t)((uint16_t)(ch16 << (bits + 1)) - ch16) + (uint16_t)addend) >> 9);
        ch16 = (uint16_t)((uint16_t)(ch16 << 1) | p);
        color[ch] = ch16;
    }
}

static void unquantize(uint16_t *color, int bits, int channels) /* :853-861 */
{
    for (int ch = 0; ch < channels; ch++)
    {
        uint16_t clr = (uint16_t)(color[ch] << (8 - bits));
        color[ch] = (uint16_t)(clr | (clr >> bits));
    }
}

static void compress_endpoints(int mode, uint16_t ep[2][4], const uint16_t p[2]) /* :862-938 (single-plane modes) */
{
    for (int j = 0; j < 2; j++)
    {
        switch (mode)
        {
        case 0: quantize_p(ep[j], 4, p[j], 3); unquantize(ep[j], 5, 3); ep[j][3] = 255; break;
        case 1: quantize_p(ep[j], 6, p[0], 3); unquantize(ep[j], 7, 3); ep[j][3] = 255; break;
        case 2: quantize(ep[j], 5, 3); unquantize(ep[j], 5, 3); ep[j][3] = 255; break;
        case 3: quantize_p(ep[j], 7, p[j], 3); ep[j][3] = 255; break;
        case 6: quantize_p(ep[j], 7, p[j], 4); break;
        case 7: quantize_p(ep[j], 5, p[j], 4); unquantize(ep[j], 6, 4); break;
        default: break;
        }
    }
}

/* g_modes, ConvectionKernels_BC67.cpp:108-119 */
enum { PBit_PerEndpoint, PBit_PerSubset, PBit_None };
enum { Alpha_Combined, Alpha_Separate, Alpha_None };
typedef struct { int pBitMode, alphaMode, rgbBits, alphaBits, partitionBits, numSubsets, indexBits, alphaIndexBits, hasIndexSelector; } mode_info;
static const mode_info kModes[8] =
{
    { PBit_PerEndpoint, Alpha_None, 4, 0, 4, 3, 3, 0, 0 },
    { PBit_PerSubset, Alpha_None, 6, 0, 6, 2, 3, 0, 0 },
    { PBit_None, Alpha_None, 5, 0, 6, 3, 2, 0, 0 },
    { PBit_PerEndpoint, Alpha_None, 7, 0, 6, 2, 2, 0, 0 },
    { PBit_None, Alpha_Separate, 5, 6, 0, 1, 2, 3, 1 },
    { PBit_None, Alpha_Separate, 7, 8, 0, 1, 2, 2, 0 },
    { PBit_PerEndpoint, Alpha_Combined, 7, 7, 0, 1, 4, 0, 0 },
    { PBit_PerEndpoint, Alpha_Combined, 5, 5, 6, 2, 2, 0, 0 }
};

/* BC67::WorkInfo, ConvectionKernels_BC67.cpp:59-76.  partition aliases indexSelector (union). */
typedef struct
{
    uint16_t mode;
    float error;
    uint16_t ep[3][2][4];
    uint16_t indexes[16];
    uint16_t indexes2[16];
    uint16_t partition_or_indexSelector;
    uint16_t rotation;
} work_info;

/* SinglePlaneTemporaries, ConvectionKernels_BC67.cpp:803-811 (zero-initialised, see header comment) */
typedef struct
{
    ufep_t unfinishedRGB[243];
    ufep_t unfinishedRGBA[129];
    uint16_t fragmentBestIndexes[1612];
    uint16_t shapeBestEP[243][2][4];
    float shapeBestError[243];
} single_plane_temps;

typedef struct
{
    int anyBlockHasAlpha; /* ConvectionKernels_BC67.cpp:1069 */
    int allowRGBModes;    /* :1072 */
} group_flags;

/* BC7Computer::TrySinglePlane, ConvectionKernels_BC67.cpp:1042-1662 */
static void try_single_plane(uint32_t flags, const uint16_t pixels[16][4], const float floatPixels[16][4], const float channelWeights[4],
    const oracle_bc7_plan *plan, int numRefineRounds, work_info *work, const group_flags *gf, single_plane_temps *temps)
{
    if (numRefineRounds < 1)
        numRefineRounds = 1;

    float channelWeightsSq[4];
    for (int ch = 0; ch < 4; ch++)
        channelWeightsSq[ch] = channelWeights[ch] * channelWeights[ch];

    memset(temps, 0, sizeof(*temps));

    uint16_t maxAlpha = 0, minAlpha = 255;
    int isPunchThrough = 1;
    for (int px = 0; px < 16; px++)
    {
        uint16_t a = pixels[px][3];
        if (a > maxAlpha) maxAlpha = a;
        if (a < minAlpha) minAlpha = a;
        isPunchThrough = isPunchThrough && (a == 0 || a == 255);
    }
    int blockHasNonMaxAlpha = (minAlpha < 255);
    int blockHasNonZeroAlpha = (0 < maxAlpha);

    const int anyBlockHasAlpha = gf->anyBlockHasAlpha;
    const int allowRGBModes = gf->allowRGBModes;
    const int allowMode7 = anyBlockHasAlpha || (plan->mode7RGBPartitionEnabled != 0);

    float preWeightedPixels[16][4];
    for (int px = 0; px < 16; px++)
        for (int ch = 0; ch < 4; ch++)
            preWeightedPixels[px][ch] = (float)pixels[px][ch] * channelWeights[ch];

    uint8_t pxlist[16];

    if (allowRGBModes) /* :1085-1110 */
    {
        for (int shapeIter = 0; shapeIter < plan->rgbNumShapesToEvaluate; shapeIter++)
        {
            int shape = plan->rgbShapeList[shapeIter];
            int n = shape_pixels(shape, pxlist);
            endpoint_selector(preWeightedPixels, 0, pxlist, n, 3, channelWeights, &temps->unfinishedRGB[shape]);
        }
    }

    for (int shapeIter = 0; shapeIter < plan->rgbaNumShapesToEvaluate; shapeIter++) /* :1113-1144 */
    {
        int shape = plan->rgbaShapeList[shapeIter];
        if (anyBlockHasAlpha || !allowRGBModes)
        {
            int n = shape_pixels(shape, pxlist);
            endpoint_selector(preWeightedPixels, 0, pxlist, n, 4, channelWeights, &temps->unfinishedRGBA[shape]);
        }
        else
        {
            /* ExpandTo<4>(255), UnfinishedEndpoints.h:93-114.  unfinishedRGB[shape] may never have been
               evaluated (plans built by ConfigureBC7EncodingPlanFromFineTuningParams never list shape 0 as an
               RGB shape); the reference then reads uninitialised stack, which is zero in the oracle build. */
            ufep_t *d = &temps->unfinishedRGBA[shape];
            const ufep_t *s = &temps->unfinishedRGB[shape];
            for (int ch = 0; ch < 3; ch++)
            {
                d->base[ch] = s->base[ch];
                d->offset[ch] = s->offset[ch];
            }
            d->base[3] = 255.0f;
            d->offset[3] = 0.0f;
        }
    }

    for (uint16_t mode = 0; mode <= 7; mode++)
    {
        if (mode == 4 || mode == 5)
            continue;
        if (mode < 4 && !allowRGBModes)
            continue;
        if (mode == 7 && !allowMode7)
            continue;

        const int isRGB = (mode < 4);
        const unsigned numPartitions = 1u << kModes[mode].partitionBits;
        const int numSubsets = kModes[mode].numSubsets;
        const int indexPrec = kModes[mode].indexBits;

        int parityBitMax = 1;
        if (kModes[mode].pBitMode == PBit_PerEndpoint)
            parityBitMax = 4;
        else if (kModes[mode].pBitMode == PBit_PerSubset)
            parityBitMax = 2;

        const int numRealChannels = isRGB ? 3 : 4;

        /* shape list of the mode (:1191-1225): membership test instead of an explicit list, same ascending order */
        uint8_t inList[243];
        memset(inList, 0, sizeof(inList));
        if (numSubsets == 1)
            inList[0] = 1;
        else if (numSubsets == 2)
            for (int s = 1; s <= 128; s++) inList[s] = 1;
        else
        {
            int np = (numPartitions == 16) ? 16 : 64;
            for (int p = 0; p < np; p++)
                for (int k = 0; k < 3; k++)
                    inList[kBC7Shapes3[p * 3 + k]] = 1;
        }

        for (int slot = 0; slot < 243; slot++)
            temps->shapeBestError[slot] = FLT_MAX;

        for (int shape = 0; shape < 243; shape++)
        {
            if (!inList[shape])
                continue;

            int numTweakRounds = isRGB ? plan->seedPointsForShapeRGB[shape] : plan->seedPointsForShapeRGBA[shape];
            if (numTweakRounds == 0)
                continue;
            if (numTweakRounds > MAX_TWEAK_ROUNDS)
                numTweakRounds = MAX_TWEAK_ROUNDS;

            int shapeStart = g_shape_start[shape];
            int shapeLength = shape_pixels(shape, pxlist);

            agg_error alphaAggError; /* AggregatedError<1> */
            agg_init(&alphaAggError);
            if (isRGB && anyBlockHasAlpha)
            {
                uint16_t filledAlpha[1] = { 255 };
                for (int pxi = 0; pxi < shapeLength; pxi++)
                {
                    uint16_t original[1] = { pixels[pxlist[pxi]][3] };
                    compute_error_ldr(filledAlpha, original, 1, &alphaAggError);
                }
            }
            float alphaWeightsSq[1] = { channelWeightsSq[3] };
            float staticAlphaError = agg_finalize(&alphaAggError, 1, flags, alphaWeightsSq);

            uint16_t tweakBaseEP[MAX_TWEAK_ROUNDS][2][4];
            memset(tweakBaseEP, 0, sizeof(tweakBaseEP));
            for (int tweak = 0; tweak < numTweakRounds; tweak++)
            {
                if (isRGB)
                {
                    ufep_finish_ldr(&temps->unfinishedRGB[shape], 3, tweak, 1 << indexPrec, tweakBaseEP[tweak][0], tweakBaseEP[tweak][1]);
                    tweakBaseEP[tweak][0][3] = tweakBaseEP[tweak][1][3] = 255;
                }
                else
                    ufep_finish_ldr(&temps->unfinishedRGBA[shape], 4, tweak, 1 << indexPrec, tweakBaseEP[tweak][0], tweakBaseEP[tweak][1]);
            }

            int punchThroughInvalid[4] = { 0, 0, 0, 0 };
            for (int pIter = 0; pIter < parityBitMax; pIter++)
            {
                if ((flags & FLAG_BC7_RespectPunchThrough) && (mode == 6 || mode == 7))
                {
                    if (pIter == 0)
                        punchThroughInvalid[pIter] = (isPunchThrough && blockHasNonZeroAlpha);
                    else if (pIter == parityBitMax - 1)
                        punchThroughInvalid[pIter] = (isPunchThrough && blockHasNonMaxAlpha);
                    else
                        punchThroughInvalid[pIter] = isPunchThrough;
                }
            }

            for (int pIter = 0; pIter < parityBitMax; pIter++)
            {
                /* per-lane view of :1300-1303,1406-1414: an invalid lane never commits */
                if (punchThroughInvalid[pIter])
                    continue;

                for (int tweak = 0; tweak < numTweakRounds; tweak++)
                {
                    uint16_t p[2];
                    p[0] = (uint16_t)(pIter & 1);
                    p[1] = (uint16_t)((pIter >> 1) & 1);

                    uint16_t ep[2][4];
                    for (int epi = 0; epi < 2; epi++)
                        for (int ch = 0; ch < 4; ch++)
                            ep[epi][ch] = tweakBaseEP[tweak][epi][ch];

                    for (int refine = 0; refine < numRefineRounds; refine++)
                    {
                        compress_endpoints(mode, ep, p);

                        float shapeError = 0.0f;

                        index_selector indexSelector;
                        isel_init(&indexSelector, 4, channelWeights, ep, 1 << indexPrec);

                        ep_refiner epRefiner;
                        refiner_init(&epRefiner, 4, 1 << indexPrec, channelWeights);

                        uint16_t indexes[16];
                        agg_error aggError;
                        agg_init(&aggError);

                        for (int pxi = 0; pxi < shapeLength; pxi++)
                        {
                            int px = pxlist[pxi];
                            uint16_t reconstructed[4];
                            uint16_t index = isel_select_index_ldr(&indexSelector, 4, floatPixels[px]);
                            isel_reconstruct_bc7(&indexSelector, index, reconstructed, numRealChannels);

                            if (flags & FLAG_BC7_FastIndexing)
                                compute_error_ldr(reconstructed, pixels[px], numRealChannels, &aggError);
                            else
                            {
                                float error = compute_error_ldr_simple(flags, reconstructed, pixels[px], 4, numRealChannels, channelWeightsSq);

                                uint16_t altIndexes[2];
                                altIndexes[0] = (uint16_t)((index > 1 ? index : 1) - 1);
                                altIndexes[1] = (uint16_t)((index + 1 < (1 << indexPrec) - 1) ? index + 1 : (1 << indexPrec) - 1);

                                for (int ii = 0; ii < 2; ii++)
                                {
                                    isel_reconstruct_bc7(&indexSelector, altIndexes[ii], reconstructed, numRealChannels);
                                    float altError = compute_error_ldr_simple(flags, reconstructed, pixels[px], 4, numRealChannels, channelWeightsSq);
                                    int better = (altError < error);
                                    error = pm_min(error, altError);
                                    if (better)
                                        index = altIndexes[ii];
                                }
                                shapeError = shapeError + error;
                            }

                            if (refine != numRefineRounds - 1)
                                refiner_contribute_unweighted_pw(&epRefiner, preWeightedPixels[px], index, numRealChannels);

                            indexes[pxi] = index;
                        }

                        if (flags & FLAG_BC7_FastIndexing)
                            shapeError = agg_finalize(&aggError, 4, flags, channelWeightsSq);

                        if (isRGB)
                            shapeError = shapeError + staticAlphaError;

                        if (shapeError < temps->shapeBestError[shape])
                        {
                            temps->shapeBestError[shape] = shapeError;
                            for (int epi = 0; epi < 2; epi++)
                                for (int ch = 0; ch < numRealChannels; ch++)
                                    temps->shapeBestEP[shape][epi][ch] = ep[epi][ch];
                            for (int pxi = 0; pxi < shapeLength; pxi++)
                                temps->fragmentBestIndexes[shapeStart + pxi] = indexes[pxi];
                        }

                        if (refine != numRefineRounds - 1)
                            refiner_get_refined_endpoints_ldr(&epRefiner, 4, ep);
                    }
                }
            }
            if (flags & FLAG_BC7_TrySingleColor)
            {
                /* Flags::BC7_TrySingleColor, :1436-1570 -> TrySingleColorRGBAMultiTable, :940-1040.  The table loop accepts an entry
                   under better = AndNot(pti, Less(avgError, bestAverageError)) (:997-998) and ParallelMath::AndNot(a, b) is
                   a & ~b (ParallelMath.h:901-906).  pti is false on every path this oracle accepts (RespectPunchThrough is
                   rejected below), so no entry of the single-colour tables is ever taken and the candidate that reaches the
                   error test is the initial one (:948-961): end points and reconstruction (0, 0, 0, 255), index 0.  The tables
                   (ConvectionKernels_BC7_SingleColor.h) and the average (:1438-1452) therefore have no observable effect and
                   are not restated. */
                static const uint16_t reconstructed[4] = { 0, 0, 0, 255 };
                agg_error aggError;
                agg_init(&aggError);
                for (int pxi = 0; pxi < shapeLength; pxi++)
                    compute_error_ldr(reconstructed, pixels[pxlist[pxi]], numRealChannels, &aggError);
                float error = agg_finalize(&aggError, 4, flags, channelWeightsSq) + staticAlphaError;
                if (error < temps->shapeBestError[shape])
                {
                    temps->shapeBestError[shape] = error;
                    for (int epi = 0; epi < 2; epi++)
                        for (int ch = 0; ch < numRealChannels; ch++)
                            temps->shapeBestEP[shape][epi][ch] = reconstructed[ch];
                    for (int pxi = 0; pxi < shapeLength; pxi++)
                        temps->fragmentBestIndexes[shapeStart + pxi] = 0;
                }
            }
        }

        /* partition scan (:1573-1660).  For mode 7 the reference assigns a misspelt variable (:1593-1596), so all 64
           partitions are scanned; un-evaluated shapes hold FLT_MAX and never win. */
        uint64_t partitionsEnabledBits = 0xffffffffffffffffULL;
        switch (mode)
        {
        case 0: partitionsEnabledBits = plan->mode0PartitionEnabled; break;
        case 1: partitionsEnabledBits = plan->mode1PartitionEnabled; break;
        case 2: partitionsEnabledBits = plan->mode2PartitionEnabled; break;
        case 3: partitionsEnabledBits = plan->mode3PartitionEnabled; break;
        case 6: partitionsEnabledBits = plan->mode6Enabled ? 1 : 0; break;
        default: break;
        }

        for (uint16_t partition = 0; partition < numPartitions; partition++)
        {
            if (((partitionsEnabledBits >> partition) & 1) == 0)
                continue;

            int partitionShapes[3] = { 0, 0, 0 };
            if (numSubsets == 2)
            {
                partitionShapes[0] = kBC7Shapes2[partition * 2];
                partitionShapes[1] = kBC7Shapes2[partition * 2 + 1];
            }
            else if (numSubsets == 3)
                for (int k = 0; k < 3; k++)
                    partitionShapes[k] = kBC7Shapes3[partition * 3 + k];

            float totalError = 0.0f;
            for (int subset = 0; subset < numSubsets; subset++)
                totalError = totalError + temps->shapeBestError[partitionShapes[subset]];

            int errorBetter = (totalError < work->error);

            if (mode == 7 && anyBlockHasAlpha)
            {
                int isRGBAllowedForThisPartition = (((plan->mode7RGBPartitionEnabled >> partition) & 1) != 0);
                if (!isRGBAllowedForThisPartition)
                    errorBetter = errorBetter && blockHasNonMaxAlpha;
            }

            if (errorBetter)
            {
                for (int subset = 0; subset < numSubsets; subset++)
                {
                    int shape = partitionShapes[subset];
                    int shapeStart = g_shape_start[shape];
                    int shapeLength = shape_pixels(shape, pxlist);

                    for (int epi = 0; epi < 2; epi++)
                        for (int ch = 0; ch < 4; ch++)
                            work->ep[subset][epi][ch] = temps->shapeBestEP[shape][epi][ch];

                    for (int pxi = 0; pxi < shapeLength; pxi++)
                        work->indexes[pxlist[pxi]] = temps->fragmentBestIndexes[shapeStart + pxi];
                }
                work->error = totalError;
                work->mode = mode;
                work->partition_or_indexSelector = partition;
            }
        }
    }
}

/* BC7Computer::TweakAlpha, ConvectionKernels_BC67.cpp:815-827 */
static void tweak_alpha(const uint16_t original[2], int tweak, int range, uint16_t result[2])
{
    float tf[2];
    compute_tweak_factors(tweak, range, tf);
    float base = (float)original[0];
    float offs = (float)original[1] - base;
    result[0] = pm_round_u15(pm_clamp(base + offs * tf[0], 0.0f, 255.0f));
    result[1] = pm_round_u15(pm_clamp(base + offs * tf[1], 0.0f, 255.0f));
}

/* BC7Computer::TryDualPlane, ConvectionKernels_BC67.cpp:1664-1965 */
static void try_dual_plane(uint32_t flags, const uint16_t pixels[16][4], const float floatPixels[16][4], const float channelWeights[4],
    const oracle_bc7_plan *plan, int numRefineRounds, work_info *work)
{
    if (numRefineRounds < 1)
        numRefineRounds = 1;

    float channelWeightsSq[4];
    for (int ch = 0; ch < 4; ch++)
        channelWeightsSq[ch] = channelWeights[ch] * channelWeights[ch];

    static const uint8_t allPixels[16] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 };

    for (uint16_t mode = 4; mode <= 5; mode++)
    {
        int numSP[2] = { 0, 0 };

        for (uint16_t rotation = 0; rotation < 4; rotation++)
        {
            if (mode == 4)
            {
                numSP[0] = plan->mode4SP[rotation][0];
                numSP[1] = plan->mode4SP[rotation][1];
            }
            else
                numSP[0] = numSP[1] = plan->mode5SP[rotation];

            if (numSP[0] == 0 && numSP[1] == 0)
                continue;

            int alphaChannel = (rotation + 3) & 3;
            int redChannel = (rotation == 1) ? 3 : 0;
            int greenChannel = (rotation == 2) ? 3 : 1;
            int blueChannel = (rotation == 3) ? 3 : 2;

            uint16_t rotatedRGB[16][4];
            float floatRotatedRGB[16][4];
            for (int px = 0; px < 16; px++)
            {
                rotatedRGB[px][0] = pixels[px][redChannel];
                rotatedRGB[px][1] = pixels[px][greenChannel];
                rotatedRGB[px][2] = pixels[px][blueChannel];
                rotatedRGB[px][3] = 0;
                for (int ch = 0; ch < 3; ch++)
                    floatRotatedRGB[px][ch] = (float)rotatedRGB[px][ch];
                floatRotatedRGB[px][3] = 0.0f;
            }

            uint16_t maxIndexSelector = (mode == 4) ? 2 : 1;

            float rotatedRGBWeights[3] = { channelWeights[redChannel], channelWeights[greenChannel], channelWeights[blueChannel] };
            float rotatedRGBWeightsSq[3] = { channelWeightsSq[redChannel], channelWeightsSq[greenChannel], channelWeightsSq[blueChannel] };
            float rotatedAlphaWeightSq[1] = { channelWeightsSq[alphaChannel] };
            float uniformWeight[1] = { 1.0f };

            float preWeightedRotatedRGB[16][4];
            for (int px = 0; px < 16; px++)
            {
                for (int ch = 0; ch < 3; ch++)
                    preWeightedRotatedRGB[px][ch] = (float)rotatedRGB[px][ch] * rotatedRGBWeights[ch];
                preWeightedRotatedRGB[px][3] = 0.0f;
            }

            for (uint16_t indexSelector = 0; indexSelector < maxIndexSelector; indexSelector++)
            {
                int numTweakRounds = numSP[indexSelector];
                if (numTweakRounds <= 0)
                    continue;
                if (numTweakRounds > MAX_TWEAK_ROUNDS)
                    numTweakRounds = MAX_TWEAK_ROUNDS;

                ufep_t unfinishedRGB;
                endpoint_selector(preWeightedRotatedRGB, 0, allPixels, 16, 3, rotatedRGBWeights, &unfinishedRGB);

                uint16_t alphaRange[2];
                alphaRange[0] = alphaRange[1] = pixels[0][alphaChannel];
                for (int px = 1; px < 16; px++)
                {
                    if (pixels[px][alphaChannel] < alphaRange[0]) alphaRange[0] = pixels[px][alphaChannel];
                    if (pixels[px][alphaChannel] > alphaRange[1]) alphaRange[1] = pixels[px][alphaChannel];
                }

                int rgbPrec, alphaPrec;
                if (mode == 4)
                {
                    rgbPrec = indexSelector ? 3 : 2;
                    alphaPrec = indexSelector ? 2 : 3;
                }
                else
                    rgbPrec = alphaPrec = 2;

                float bestRGBError = FLT_MAX, bestAlphaError = FLT_MAX;
                uint16_t bestRGBIndexes[16], bestAlphaIndexes[16];
                uint16_t bestEP[2][4];
                memset(bestEP, 0, sizeof(bestEP));
                for (int px = 0; px < 16; px++)
                    bestRGBIndexes[px] = bestAlphaIndexes[px] = 0;

                for (int tweak = 0; tweak < numTweakRounds; tweak++)
                {
                    uint16_t rgbEP[2][4];
                    uint16_t alphaEP[2];
                    memset(rgbEP, 0, sizeof(rgbEP));

                    ufep_finish_ldr(&unfinishedRGB, 3, tweak, 1 << rgbPrec, rgbEP[0], rgbEP[1]);
                    tweak_alpha(alphaRange, tweak, 1 << alphaPrec, alphaEP);

                    for (int refine = 0; refine < numRefineRounds; refine++)
                    {
                        /* CompressEndpoints4 / CompressEndpoints5, :896-920 */
                        for (int j = 0; j < 2; j++)
                        {
                            if (mode == 4)
                            {
                                quantize(rgbEP[j], 5, 3);
                                unquantize(rgbEP[j], 5, 3);
                                quantize(alphaEP + j, 6, 1);
                                unquantize(alphaEP + j, 6, 1);
                            }
                            else
                            {
                                quantize(rgbEP[j], 7, 3);
                                unquantize(rgbEP[j], 7, 3);
                            }
                        }

                        index_selector alphaIndexSelector, rgbIndexSelector;
                        {
                            uint16_t alphaEPTemp[2][4] = { { alphaEP[0], 0, 0, 0 }, { alphaEP[1], 0, 0, 0 } };
                            isel_init(&alphaIndexSelector, 1, uniformWeight, alphaEPTemp, 1 << alphaPrec);
                        }
                        isel_init(&rgbIndexSelector, 3, rotatedRGBWeights, rgbEP, 1 << rgbPrec);

                        ep_refiner rgbRefiner, alphaRefiner;
                        refiner_init(&rgbRefiner, 3, 1 << rgbPrec, rotatedRGBWeights);
                        refiner_init(&alphaRefiner, 1, 1 << alphaPrec, uniformWeight);

                        float errorRGB = 0.0f, errorA = 0.0f;
                        uint16_t rgbIndexes[16], alphaIndexes[16];
                        agg_error rgbAggError, alphaAggError;
                        agg_init(&rgbAggError);
                        agg_init(&alphaAggError);

                        for (int px = 0; px < 16; px++)
                        {
                            uint16_t rgbIndex = isel_select_index_ldr(&rgbIndexSelector, 3, floatRotatedRGB[px]);
                            uint16_t alphaIndex = isel_select_index_ldr(&alphaIndexSelector, 1, floatPixels[px] + alphaChannel);

                            uint16_t reconstructedRGB[4], reconstructedAlpha[1];
                            isel_reconstruct_bc7(&rgbIndexSelector, rgbIndex, reconstructedRGB, 3);
                            isel_reconstruct_bc7(&alphaIndexSelector, alphaIndex, reconstructedAlpha, 1);

                            if (flags & FLAG_BC7_FastIndexing)
                            {
                                compute_error_ldr(reconstructedRGB, rotatedRGB[px], 3, &rgbAggError);
                                compute_error_ldr(reconstructedAlpha, pixels[px] + alphaChannel, 1, &alphaAggError);
                            }
                            else
                            {
                                float rgbError = compute_error_ldr_simple(flags, reconstructedRGB, rotatedRGB[px], 3, 3, rotatedRGBWeightsSq);
                                float alphaError = compute_error_ldr_simple(flags, reconstructedAlpha, pixels[px] + alphaChannel, 1, 1, rotatedAlphaWeightSq);

                                uint16_t altRGBIndexes[2], altAlphaIndexes[2];
                                altRGBIndexes[0] = (uint16_t)((rgbIndex > 1 ? rgbIndex : 1) - 1);
                                altRGBIndexes[1] = (uint16_t)((rgbIndex + 1 < (1 << rgbPrec) - 1) ? rgbIndex + 1 : (1 << rgbPrec) - 1);
                                altAlphaIndexes[0] = (uint16_t)((alphaIndex > 1 ? alphaIndex : 1) - 1);
                                altAlphaIndexes[1] = (uint16_t)((alphaIndex + 1 < (1 << alphaPrec) - 1) ? alphaIndex + 1 : (1 << alphaPrec) - 1);

                                for (int ii = 0; ii < 2; ii++)
                                {
                                    isel_reconstruct_bc7(&rgbIndexSelector, altRGBIndexes[ii], reconstructedRGB, 3);
                                    isel_reconstruct_bc7(&alphaIndexSelector, altAlphaIndexes[ii], reconstructedAlpha, 1);

                                    float altRGBError = compute_error_ldr_simple(flags, reconstructedRGB, rotatedRGB[px], 3, 3, rotatedRGBWeightsSq);
                                    float altAlphaError = compute_error_ldr_simple(flags, reconstructedAlpha, pixels[px] + alphaChannel, 1, 1, rotatedAlphaWeightSq);

                                    int rgbBetter = (altRGBError < rgbError);
                                    int alphaBetter = (altAlphaError < alphaError);

                                    rgbError = pm_min(altRGBError, rgbError);
                                    alphaError = pm_min(altAlphaError, alphaError);

                                    if (rgbBetter) rgbIndex = altRGBIndexes[ii];
                                    if (alphaBetter) alphaIndex = altAlphaIndexes[ii];
                                }

                                errorRGB = errorRGB + rgbError;
                                errorA = errorA + alphaError;
                            }

                            if (refine != numRefineRounds - 1)
                            {
                                refiner_contribute_unweighted_pw(&rgbRefiner, preWeightedRotatedRGB[px], rgbIndex, 3);
                                refiner_contribute_unweighted_pw(&alphaRefiner, floatPixels[px] + alphaChannel, alphaIndex, 1);
                            }

                            if (flags & FLAG_BC7_FastIndexing)
                            {
                                errorRGB = agg_finalize(&rgbAggError, 3, flags, rotatedRGBWeightsSq);
                                errorA = agg_finalize(&alphaAggError, 1, flags, rotatedAlphaWeightSq);
                            }

                            rgbIndexes[px] = rgbIndex;
                            alphaIndexes[px] = alphaIndex;
                        }

                        if (errorRGB < bestRGBError)
                        {
                            bestRGBError = pm_min(errorRGB, bestRGBError);
                            for (int px = 0; px < 16; px++)
                                bestRGBIndexes[px] = rgbIndexes[px];
                            for (int ep = 0; ep < 2; ep++)
                                for (int ch = 0; ch < 3; ch++)
                                    bestEP[ep][ch] = rgbEP[ep][ch];
                        }

                        if (errorA < bestAlphaError)
                        {
                            bestAlphaError = pm_min(errorA, bestAlphaError);
                            for (int px = 0; px < 16; px++)
                                bestAlphaIndexes[px] = alphaIndexes[px];
                            for (int ep = 0; ep < 2; ep++)
                                bestEP[ep][3] = alphaEP[ep];
                        }

                        if (refine != numRefineRounds - 1)
                        {
                            refiner_get_refined_endpoints_ldr(&rgbRefiner, 3, rgbEP);

                            uint16_t alphaEPTemp[2][4];
                            refiner_get_refined_endpoints_ldr(&alphaRefiner, 1, alphaEPTemp);
                            for (int i = 0; i < 2; i++)
                                alphaEP[i] = alphaEPTemp[i][0];
                        }
                    }
                }

                float combinedError = bestRGBError + bestAlphaError;
                int errorBetter = (combinedError < work->error);
                work->error = pm_min(combinedError, work->error);

                if (errorBetter)
                {
                    work->mode = mode;
                    work->rotation = rotation;
                    work->partition_or_indexSelector = indexSelector;
                    for (int px = 0; px < 16; px++)
                    {
                        work->indexes[px] = indexSelector ? bestAlphaIndexes[px] : bestRGBIndexes[px];
                        work->indexes2[px] = indexSelector ? bestRGBIndexes[px] : bestAlphaIndexes[px];
                    }
                    for (int ep = 0; ep < 2; ep++)
                        for (int ch = 0; ch < 4; ch++)
                            work->ep[0][ep][ch] = bestEP[ep][ch];
                }
            }
        }
    }
}

/* PackingVector, ConvectionKernels_BC67.cpp:652-698 */
typedef struct { uint32_t v[5]; int offset; } packing_vector;

static void pv_pack(packing_vector *pv, uint16_t value, int bits)
{
    int vOffset = pv->offset >> 5;
    int bitOffset = pv->offset & 0x1f;
    pv->v[vOffset] |= ((uint32_t)value << bitOffset);
    int overflowBits = bitOffset + bits - 32;
    if (overflowBits > 0)
        pv->v[vOffset + 1] |= ((uint32_t)value >> (bits - overflowBits));
    pv->offset += bits;
}

static void swap_u16(uint16_t *a, uint16_t *b) { uint16_t t = *a; *a = *b; *b = t; }

/* per-block tail of BC7Computer::Pack, ConvectionKernels_BC67.cpp:2003-2203 */
static void pack_block(const work_info *work, uint8_t *output)
{
    packing_vector pv;
    memset(&pv, 0, sizeof(pv));

    uint16_t mode = work->mode;
    uint16_t partition = work->partition_or_indexSelector;
    uint16_t indexSelector = work->partition_or_indexSelector;
    const mode_info *mi = &kModes[mode];

    uint16_t indexes[16], indexes2[16], endPoints[3][2][4];
    memcpy(indexes, work->indexes, sizeof(indexes));
    memcpy(indexes2, work->indexes2, sizeof(indexes2));
    memcpy(endPoints, work->ep, sizeof(endPoints));

    int fixups[3] = { 0, 0, 0 };

    if (mi->alphaMode == Alpha_Separate)
    {
        int flipRGB = ((indexes[0] & (1 << (mi->indexBits - 1))) != 0);
        int flipAlpha = ((indexes2[0] & (1 << (mi->alphaIndexBits - 1))) != 0);

        if (flipRGB)
        {
            uint16_t highIndex = (uint16_t)((1 << mi->indexBits) - 1);
            for (int px = 0; px < 16; px++)
                indexes[px] = (uint16_t)(highIndex - indexes[px]);
        }
        if (flipAlpha)
        {
            uint16_t highIndex = (uint16_t)((1 << mi->alphaIndexBits) - 1);
            for (int px = 0; px < 16; px++)
                indexes2[px] = (uint16_t)(highIndex - indexes2[px]);
        }
        if (indexSelector)
        {
            int t = flipRGB; flipRGB = flipAlpha; flipAlpha = t;
        }
        if (flipRGB)
            for (int ch = 0; ch < 3; ch++)
                swap_u16(&endPoints[0][0][ch], &endPoints[0][1][ch]);
        if (flipAlpha)
            swap_u16(&endPoints[0][0][3], &endPoints[0][1][3]);
    }
    else
    {
        if (mi->numSubsets == 2)
            fixups[1] = kBC7Fixup2[partition];
        else if (mi->numSubsets == 3)
        {
            fixups[1] = kBC7Fixup3[partition * 2];
            fixups[2] = kBC7Fixup3[partition * 2 + 1];
        }

        int flip[3] = { 0, 0, 0 };
        for (int subset = 0; subset < mi->numSubsets; subset++)
            flip[subset] = ((indexes[fixups[subset]] & (1 << (mi->indexBits - 1))) != 0);

        if (flip[0] || flip[1] || flip[2])
        {
            uint16_t highIndex = (uint16_t)((1 << mi->indexBits) - 1);
            for (int px = 0; px < 16; px++)
            {
                int subset = 0;
                if (mi->numSubsets == 2)
                    subset = (kBC7PartitionMask2[partition] >> px) & 1;
                else if (mi->numSubsets == 3)
                    subset = (kBC7PartitionMap3[partition] >> (px * 2)) & 3;
                if (flip[subset])
                    indexes[px] = (uint16_t)(highIndex - indexes[px]);
            }
            int maxCH = (mi->alphaMode == Alpha_Combined) ? 4 : 3;
            for (int subset = 0; subset < mi->numSubsets; subset++)
                if (flip[subset])
                    for (int ch = 0; ch < maxCH; ch++)
                        swap_u16(&endPoints[subset][0][ch], &endPoints[subset][1][ch]);
        }
    }

    pv_pack(&pv, (uint8_t)(1 << mode), mode + 1);
    if (mi->partitionBits)
        pv_pack(&pv, partition, mi->partitionBits);
    if (mi->alphaMode == Alpha_Separate)
        pv_pack(&pv, work->rotation, 2);
    if (mi->hasIndexSelector)
        pv_pack(&pv, indexSelector, 1);

    for (int ch = 0; ch < 3; ch++)
        for (int subset = 0; subset < mi->numSubsets; subset++)
            for (int ep = 0; ep < 2; ep++)
                pv_pack(&pv, (uint16_t)(endPoints[subset][ep][ch] >> (8 - mi->rgbBits)), mi->rgbBits);

    if (mi->alphaMode != Alpha_None)
        for (int subset = 0; subset < mi->numSubsets; subset++)
            for (int ep = 0; ep < 2; ep++)
                pv_pack(&pv, (uint16_t)(endPoints[subset][ep][3] >> (8 - mi->alphaBits)), mi->alphaBits);

    if (mi->pBitMode == PBit_PerSubset)
    {
        for (int subset = 0; subset < mi->numSubsets; subset++)
            pv_pack(&pv, (uint16_t)((endPoints[subset][0][0] >> (7 - mi->rgbBits)) & 1), 1);
    }
    else if (mi->pBitMode == PBit_PerEndpoint)
    {
        for (int subset = 0; subset < mi->numSubsets; subset++)
            for (int ep = 0; ep < 2; ep++)
                pv_pack(&pv, (uint16_t)((endPoints[subset][ep][0] >> (7 - mi->rgbBits)) & 1), 1);
    }

    for (int px = 0; px < 16; px++)
    {
        int bits = mi->indexBits;
        if ((px == 0) || (px == fixups[1]) || (px == fixups[2]))
            bits--;
        pv_pack(&pv, indexes[px], bits);
    }

    if (mi->alphaMode == Alpha_Separate)
        for (int px = 0; px < 16; px++)
        {
            int bits = mi->alphaIndexBits;
            if (px == 0)
                bits--;
            pv_pack(&pv, indexes2[px], bits);
        }

    for (int v = 0; v < 4; v++)
        for (int b = 0; b < 4; b++)
            output[v * 4 + b] = (uint8_t)((pv.v[v] >> (b * 8)) & 0xff);
}

/* ---- entry points ---- */

/* Kernels::EncodeBC7 + BC7Computer::Pack for nBlocks = 8*k blocks (ConvectionKernels_API.cpp:41-54,
 * ConvectionKernels_BC67.cpp:1975-2001).  Returns 0, -1 bad argument, -2 unsupported flag. */
int cvtt_oracle_encode_bc7(const uint8_t *blocks, size_t nBlocks, uint8_t *out, const oracle_options *options, const oracle_bc7_plan *plan)
{
    if (!blocks || !out || !options || !plan || (nBlocks % 8) != 0)
        return -1;
    /* BC7_RespectPunchThrough is not restated: the reference masks commits with AndNot(punchThroughInvalid, better) = invalid & ~better
       (ConvectionKernels_BC67.cpp:1411, ParallelMath.h:898-903), guarded by AnySet(better) over the 8 lanes (:1406),
       so a block's result depends on its neighbours' per-trial errors; that needs a lock-step 8-lane model. */
    if (options->flags & FLAG_BC7_RespectPunchThrough)
        return -2;

    init_tables();

    float channelWeights[4]; /* Util::FillWeights, ConvectionKernels_Util.cpp:62-73 */
    if (options->flags & FLAG_Uniform)
        channelWeights[0] = channelWeights[1] = channelWeights[2] = channelWeights[3] = 1.0f;
    else
    {
        channelWeights[0] = options->redWeight;
        channelWeights[1] = options->greenWeight;
        channelWeights[2] = options->blueWeight;
        channelWeights[3] = options->alphaWeight;
    }

    static __thread single_plane_temps temps;

    for (size_t g = 0; g < nBlocks / 8; g++)
    {
        const uint8_t *grp = blocks + g * 8 * 64;
        group_flags gf = { 0, 0 };
        for (int b = 0; b < 8; b++)
        {
            int minAlpha = 255;
            for (int px = 0; px < 16; px++)
            {
                int a = grp[b * 64 + px * 4 + 3];
                if (a < minAlpha) minAlpha = a;
            }
            if (minAlpha < 255) gf.anyBlockHasAlpha = 1;
            if (250 < minAlpha) gf.allowRGBModes = 1;
        }

        for (int b = 0; b < 8; b++)
        {
            uint16_t pixels[16][4];
            float floatPixels[16][4];
            for (int px = 0; px < 16; px++)
                for (int ch = 0; ch < 4; ch++)
                {
                    pixels[px][ch] = grp[b * 64 + px * 4 + ch];
                    floatPixels[px][ch] = (float)pixels[px][ch];
                }

            work_info work;
            memset(&work, 0, sizeof(work));
            work.error = FLT_MAX;

            try_single_plane(options->flags, pixels, floatPixels, channelWeights, plan, options->refineRoundsBC7, &work, &gf, &temps);
            try_dual_plane(options->flags, pixels, floatPixels, channelWeights, plan, options->refineRoundsBC7, &work);
            pack_block(&work, out + (g * 8 + b) * 16);
        }
    }
    return 0;
}

/* ---- host-side plan configuration: Kernels::ConfigureBC7EncodingPlanFromFineTuningParams / FromQuality,
 *      ConvectionKernels_BC67.cpp:3291-3483 ---- */
typedef struct
{
    uint8_t mode0SP[16], mode1SP[64], mode2SP[64], mode3SP[64];
    uint8_t mode4SP[4][2];
    uint8_t mode5SP[4];
    uint8_t mode6SP;
    uint8_t mode7SP[64];
} oracle_bc7_finetune;

static uint8_t max_u8(uint8_t a, uint8_t b) { return a > b ? a : b; }

int cvtt_oracle_plan_from_finetune(oracle_bc7_plan *plan, const oracle_bc7_finetune *params)
{
    memset(plan, 0, sizeof(*plan));

    for (int partition = 0; partition < 16; partition++)
    {
        uint8_t sp = params->mode0SP[partition];
        if (sp == 0) continue;
        plan->mode0PartitionEnabled |= (uint16_t)(1u << partition);
        for (int subset = 0; subset < 3; subset++)
        {
            int shape = kBC7Shapes3[partition * 3 + subset];
            plan->seedPointsForShapeRGB[shape] = max_u8(plan->seedPointsForShapeRGB[shape], sp);
        }
    }
    for (int partition = 0; partition < 64; partition++)
    {
        uint8_t sp = params->mode1SP[partition];
        if (sp == 0) continue;
        plan->mode1PartitionEnabled |= (uint64_t)1 << partition;
        for (int subset = 0; subset < 2; subset++)
        {
            int shape = kBC7Shapes2[partition * 2 + subset];
            plan->seedPointsForShapeRGB[shape] = max_u8(plan->seedPointsForShapeRGB[shape], sp);
        }
    }
    for (int partition = 0; partition < 64; partition++)
    {
        uint8_t sp = params->mode2SP[partition];
        if (sp == 0) continue;
        plan->mode2PartitionEnabled |= (uint64_t)1 << partition;
        for (int subset = 0; subset < 3; subset++)
        {
            int shape = kBC7Shapes3[partition * 3 + subset];
            plan->seedPointsForShapeRGB[shape] = max_u8(plan->seedPointsForShapeRGB[shape], sp);
        }
    }
    for (int partition = 0; partition < 64; partition++)
    {
        uint8_t sp = params->mode3SP[partition];
        if (sp == 0) continue;
        plan->mode3PartitionEnabled |= (uint64_t)1 << partition;
        for (int subset = 0; subset < 2; subset++)
        {
            int shape = kBC7Shapes2[partition * 2 + subset];
            plan->seedPointsForShapeRGB[shape] = max_u8(plan->seedPointsForShapeRGB[shape], sp);
        }
    }
    for (int rotation = 0; rotation < 4; rotation++)
    {
        for (int indexMode = 0; indexMode < 2; indexMode++)
            plan->mode4SP[rotation][indexMode] = params->mode4SP[rotation][indexMode];
        plan->mode5SP[rotation] = params->mode5SP[rotation];
    }
    if (params->mode6SP != 0)
    {
        plan->mode6Enabled = 1;
        plan->seedPointsForShapeRGBA[0] = max_u8(plan->seedPointsForShapeRGBA[0], params->mode6SP);
    }
    for (int partition = 0; partition < 64; partition++)
    {
        uint8_t sp = params->mode7SP[partition];
        if (sp == 0) continue;
        plan->mode7RGBAPartitionEnabled |= (uint64_t)1 << partition;
        for (int subset = 0; subset < 2; subset++)
        {
            int shape = kBC7Shapes2[partition * 2 + subset];
            plan->seedPointsForShapeRGBA[shape] = max_u8(plan->seedPointsForShapeRGBA[shape], sp);
        }
    }
    for (int i = 0; i < 243; i++)
        if (plan->seedPointsForShapeRGB[i] > 0)
            plan->rgbShapeList[plan->rgbNumShapesToEvaluate++] = (uint8_t)i;
    for (int i = 0; i < 129; i++)
        if (plan->seedPointsForShapeRGBA[i] > 0)
            plan->rgbaShapeList[plan->rgbaNumShapesToEvaluate++] = (uint8_t)i;

    plan->mode7RGBPartitionEnabled = (plan->mode7RGBAPartitionEnabled & ~plan->mode3PartitionEnabled);
    return 1;
}

void cvtt_oracle_plan_from_quality(oracle_bc7_plan *plan, int quality)
{
    if (quality < 1) quality = 1;
    else if (quality > 100) quality = 100;

    const int counts[2] = { CVTT_BC7_NUM_PRIO_RGB * quality / 100, CVTT_BC7_NUM_PRIO_RGBA * quality / 100 };
    const unsigned short *lists[2] = { kBC7PrioRGB, kBC7PrioRGBA };

    oracle_bc7_finetune ft;
    memset(&ft, 0, sizeof(ft));

    for (int li = 0; li < 2; li++)
        for (int i = 0; i < counts[li]; i++)
        {
            unsigned code = lists[li][i];
            uint8_t sp = (uint8_t)(((code >> 9) & 3) + 1);
            int mode = (code >> 6) & 7;
            int partition = code & 63, rotation = code & 3, isel = (code >> 2) & 1;
            switch (mode)
            {
            case 0: ft.mode0SP[partition] = sp; break;
            case 1: ft.mode1SP[partition] = sp; break;
            case 2: ft.mode2SP[partition] = sp; break;
            case 3: ft.mode3SP[partition] = sp; break;
            case 4: ft.mode4SP[rotation][isel] = sp; break;
            case 5: ft.mode5SP[rotation] = sp; break;
            case 6: ft.mode6SP = sp; break;
            case 7: ft.mode7SP[partition] = sp; break;
            }
        }
    cvtt_oracle_plan_from_finetune(plan, &ft);
}

size_t cvtt_oracle_sizeof_plan(void) { return sizeof(oracle_bc7_plan); }
size_t cvtt_oracle_sizeof_options(void) { return sizeof(oracle_options); }
