// Diagnostic: how fast can the host CPUs expand a sparse ray transport into dense float32 rows?  Per env: a 960-byte
// template row (the no-hit values) + ~64 scattered hit values.  Threads x {memcpy, non-temporal stores}.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <emmintrin.h>
static void fill_rows(float* dst, const float* tmpl, const float* hits, const unsigned* mask, int env0, int env1, int W, bool nt) {
    for (int e = env0; e < env1; e++) {
        float* row = dst + (size_t)e * W;
        if (nt) {
            for (int k = 0; k < W; k += 4) _mm_stream_si128((__m128i*)(row + k), _mm_load_si128((const __m128i*)(tmpl + k)));
        } else {
            memcpy(row, tmpl, sizeof(float) * W);
        }
        const unsigned* m = mask + (size_t)e * 8;
        const float* h = hits + (size_t)e * 64;
        int j = 0;
        for (int w = 0; w < 8; w++) {
            unsigned b = m[w];
            while (b) { int k = __builtin_ctz(b); b &= b - 1; row[w * 32 + k] = h[j++ & 63]; }
        }
    }
}
int main(int argc, char** argv) {
    const int n = 65536, W = 240;
    float* dst = (float*)aligned_alloc(4096, sizeof(float) * (size_t)n * W);
    float* tmpl = (float*)aligned_alloc(64, sizeof(float) * W);
    std::vector<float> hits((size_t)n * 64, 1.f);
    std::vector<unsigned> mask((size_t)n * 8);
    for (int k = 0; k < W; k++) tmpl[k] = 200.f;
    unsigned s = 12345;
    for (size_t i = 0; i < mask.size(); i++) { unsigned v = 0; for (int b = 0; b < 32; b++) { s = s * 1664525u + 1013904223u; if ((s >> 24) < 68 && (i % 8) * 32 + b < (size_t)W) v |= 1u << b; } mask[i] = v; }
    memset(dst, 0, sizeof(float) * (size_t)n * W);
    for (int nt = 0; nt < 2; nt++)
        for (int T : {1, 2, 4, 8, 16, 32}) {
            if (T > (int)std::thread::hardware_concurrency()) continue;
            double best = 1e9;
            for (int rep = 0; rep < 6; rep++) {
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                for (int t = 0; t < T; t++) th.emplace_back(fill_rows, dst, tmpl, hits.data(), mask.data(), (int)((long long)n * t / T), (int)((long long)n * (t + 1) / T), W, nt != 0);
                for (auto& x : th) x.join();
                double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                if (ms < best) best = ms;
            }
            printf("%s threads %2d: %.3f ms per 65536 envs (%.1f GB/s of output)\n", nt ? "nontemporal" : "memcpy     ", T, best, n * W * 4 / best / 1e6);
        }
    printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
    return 0;
}
