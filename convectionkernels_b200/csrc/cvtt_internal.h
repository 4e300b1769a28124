// Host-side plumbing shared by the translation units of libcvtt_b200.so: error reporting, the per-device context and the
// launch entry points each kernel TU exports (bc7_kernels.cu, bc6h_kernels.cu, etc_kernels.cu, s3tc_kernels.cu).  The TUs
// are independent (own kernels, own __constant__ tables), so the library needs no relocatable device code and the TUs
// compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include <stdint.h>
#include <string.h>

#include "../../include/cvtt_b200.h"
#include "cvtt_common.cuh"

namespace cvttb200
{
    extern std::atomic<uint64_t> g_launches;          // kernels launched by this library (cvttb200_launch_count)

    int fail(int code, const std::string &msg);
    int fail_cuda(cudaError_t e, const char *what);

#define CVTT_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ::cvttb200::fail_cuda(e_, #call); } while (0)

    struct PlanCacheEntry
    {
        BC7PlanPOD plan;
        int form;                                     // which command-stream form (bc7_compile_plan: plain, pair, split)
        uint32_t *dCmds;
    };

    // One per device, created on first use and kept until cvttb200_shutdown.  Calls from several host threads share it; see
    // "Host state" in cvtt_b200.cu for what each mutex covers.
    struct DeviceContext
    {
        int device = -1;
        bool ready = false;
        int numSMs = 0;
        cudaMemPool_t pool = nullptr;                 // the library's own stream-ordered pool: staging buffers, kernel scratch
        std::mutex planMutex;                         // plans, setupStream; held while a launch that reads a cached plan is enqueued
        std::vector<PlanCacheEntry> plans;
        cudaStream_t setupStream = nullptr;           // uploads of compiled plans (synchronised before the plan is published)
        std::mutex pipeMutex;                         // pipeStream
        cudaStream_t pipeStream[2] = { nullptr, nullptr };   // host-buffer calls of the fast formats: chunks alternate between two streams
        cudaStream_t multiStream = nullptr;           // this device's stream of cvttb200_encode_multi (serialised by its own mutex)
        std::mutex streamMutex;                       // idleStreams
        std::vector<cudaStream_t> idleStreams;        // non-blocking streams for host-buffer calls that name no stream (one per call in flight)
    };

    // stream-ordered allocation from ctx.pool (freed with cudaFreeAsync on the same stream)
    int pool_alloc(DeviceContext &ctx, void **p, size_t bytes, cudaStream_t stream);

    // Per-TU device set-up (constant tables, kernel attributes) for the current device, and the launches.  All return a
    // CVTTB200_* code; the context's device is current; launches may run concurrently on several host threads.  rcpN is the
    // host's _mm_rcp_ps table (cvttb200_set_rcp_table).
    int bc7_device_setup();
    int launch_bc7(DeviceContext &ctx, const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, const BC7PlanPOD &plan, const float *rcpN, cudaStream_t stream);
    int bc7_selftest_div(uint64_t samples, uint64_t seed, uint64_t *mismatches);

    int bc6h_device_setup();
    int launch_bc6h(DeviceContext &ctx, const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, bool isSigned, const float *rcpN, cudaStream_t stream);

    int etc_device_setup();
    // allocOptions: the Options AllocETC2Data was called with (the reference fixes the chroma side axes there, ETC.cpp:3117-3145)
    int launch_etc(DeviceContext &ctx, int format, const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, const OptionsPOD &allocOptions, cudaStream_t stream);

    int launch_s3tc(int format, const void *dIn, size_t nBlocks, void *dOut, const OptionsPOD &options, const float *rcpN, cudaStream_t stream);
}
