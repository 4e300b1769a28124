// A reference-style caller (cf. reference etc2packer/etc2packer.cpp:215-282): packs 8 horizontally adjacent 4x4 blocks per
// call and encodes them with cvtt::Kernels::EncodeBC7, then encodes the same image with the whole-image entry point and
// checks both agree.  Prints a 64-bit FNV hash of the output so a harness can compare it with the oracle's.
// usage: dropin_main <width> <height> <quality>
#include <thread>
#include <vector>
#include <stdint.h>
#include <string.h>
#include "cvtt_b200_dropin.h"

int main(int argc, char **argv)
{
    const int w = argc > 1 ? atoi(argv[1]) : 64, h = argc > 2 ? atoi(argv[2]) : 64, quality = argc > 3 ? atoi(argv[3]) : 100;
    std::vector<uint8_t> image((size_t)w * h * 4);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            uint8_t *p = &image[((size_t)y * w + x) * 4];
            p[0] = (uint8_t)x; p[1] = (uint8_t)y; p[2] = (uint8_t)((x + y) / 2); p[3] = (uint8_t)(((x / 32) & 1) ? 255 : 255 - y);
        }

    cvtt::Options options;
    cvtt::BC7EncodingPlan plan;
    cvtt::Kernels::ConfigureBC7EncodingPlanFromQuality(plan, quality);

    const int bw = w / 4, bh = h / 4;
    std::vector<cvtt::PixelBlockU8> blocks((size_t)bw * bh);
    for (int by = 0; by < bh; by++)
        for (int bx = 0; bx < bw; bx++)
            for (int py = 0; py < 4; py++)
                memcpy(blocks[(size_t)by * bw + bx].m_pixels[py * 4], &image[(((size_t)by * 4 + py) * w + bx * 4) * 4], 16);

    std::vector<uint8_t> perCall(blocks.size() * 16), whole(blocks.size() * 16);
    for (size_t b = 0; b < blocks.size(); b += cvtt::NumParallelBlocks)
        cvtt::Kernels::EncodeBC7(&perCall[b * 16], &blocks[b], options, plan);
    cvtt::Kernels::B200::EncodeBC7(whole.data(), blocks.data(), blocks.size(), options, plan);

    if (memcmp(perCall.data(), whole.data(), whole.size()) != 0)
    {
        fprintf(stderr, "8-block calls and whole-image call disagree\n");
        return 2;
    }
    // the other reference entry points, 8 blocks per call vs one whole-image call
    {
        static void *(*allocShim)(void *, size_t) = [](void *, size_t size) -> void * { return malloc(size); };
        static void (*freeShim)(void *, void *, size_t) = [](void *, void *ptr, size_t) { free(ptr); };
        cvtt::ETC2CompressionData *etc2 = cvtt::Kernels::AllocETC2Data(*allocShim, NULL, options);
        std::vector<uint8_t> a(blocks.size() * 16), b(blocks.size() * 16);
        for (size_t i = 0; i < blocks.size(); i += cvtt::NumParallelBlocks)
            cvtt::Kernels::EncodeETC2RGBA(&a[i * 16], &blocks[i], options, etc2);
        cvtt::Kernels::B200::Encode(CVTTB200_ETC2_RGBA, b.data(), blocks.data(), blocks.size(), options);
        cvtt::Kernels::ReleaseETC2Data(etc2, *freeShim);
        if (a != b)
        {
            fprintf(stderr, "ETC2 RGBA: 8-block calls and whole-image call disagree\n");
            return 3;
        }
        for (size_t i = 0; i < blocks.size(); i += cvtt::NumParallelBlocks)
            cvtt::Kernels::EncodeBC3(&a[i * 16], &blocks[i], options);
        cvtt::Kernels::B200::Encode(CVTTB200_BC3, b.data(), blocks.data(), blocks.size(), options);
        if (a != b)
        {
            fprintf(stderr, "BC3: 8-block calls and whole-image call disagree\n");
            return 4;
        }
        uint64_t h2 = 1469598103934665603ull;
        for (size_t i = 0; i < b.size(); i++)
            h2 = (h2 ^ b[i]) * 1099511628211ull;
        printf("BC3 fnv64 %016llx\n", (unsigned long long)h2);

        // AllocETC2Data with options of its own (the chroma side axes come from these), EncodeETC2 with others
        cvtt::Options encOptions, allocOptions;
        encOptions.redWeight = 1.0f; encOptions.greenWeight = 0.5f; encOptions.blueWeight = 0.25f;
        allocOptions.redWeight = 0.1f; allocOptions.greenWeight = 1.0f; allocOptions.blueWeight = 0.7f;
        cvtt::ETC2CompressionData *etc2b = cvtt::Kernels::AllocETC2Data(*allocShim, NULL, allocOptions);
        std::vector<uint8_t> c(blocks.size() * 8);
        for (size_t i = 0; i < blocks.size(); i += cvtt::NumParallelBlocks)
            cvtt::Kernels::EncodeETC2(&c[i * 8], &blocks[i], encOptions, etc2b);
        cvtt::Kernels::ReleaseETC2Data(etc2b, *freeShim);
        uint64_t h3 = 1469598103934665603ull;
        for (size_t i = 0; i < c.size(); i++)
            h3 = (h3 ^ c[i]) * 1099511628211ull;
        printf("ETC2 alloc-options fnv64 %016llx\n", (unsigned long long)h3);
    }
    // the reference may be called from any number of threads (README.md:57): four threads encode quarters of the image with
    // 8-block calls at the same time
    {
        std::vector<uint8_t> threaded(blocks.size() * 16);
        std::vector<std::thread> pool;
        const size_t groups = blocks.size() / cvtt::NumParallelBlocks;
        for (int t = 0; t < 4; t++)
            pool.emplace_back([&, t]() {
                for (size_t g = groups * t / 4; g < groups * (t + 1) / 4; g++)
                    cvtt::Kernels::EncodeBC7(&threaded[g * 8 * 16], &blocks[g * 8], options, plan);
            });
        for (size_t t = 0; t < pool.size(); t++)
            pool[t].join();
        if (memcmp(threaded.data(), whole.data(), whole.size()) != 0)
        {
            fprintf(stderr, "four caller threads disagree with the whole-image call\n");
            return 5;
        }
        printf("4 caller threads OK\n");
    }

    uint64_t hash = 1469598103934665603ull;
    for (size_t i = 0; i < whole.size(); i++)
        hash = (hash ^ whole[i]) * 1099511628211ull;
    printf("%zu blocks, fnv64 %016llx\n", blocks.size(), (unsigned long long)hash);
    return 0;
}
