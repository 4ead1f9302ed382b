// host_probe.c -- how fast can the HOST side of the window tensor be written?
//
// The (N, 11, C) float32 input tensor is 11x redundant: row k of an utterance is the 11*C
// floats that start at decimated frame k.  Shipping decimated frames over PCIe (0.68 GB per
// corpus) and expanding them on the host turns the 7.5 GB device->host copy into a host
// memory-write problem.  This probe measures exactly that access pattern -- 5632-byte rows
// read from a sliding 512-byte-stride source, written back to back -- by thread count, with
// plain memcpy and with non-temporal stores, on warm and on never-touched destination pages.
//
//   gcc -O2 -pthread -o tools/host_probe tools/host_probe.c && tools/host_probe [dst_GB]
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

static double now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

enum { ROW_FLOATS = 11 * 128, FRAME_FLOATS = 128 };

typedef struct {
    const float* src;
    float* dst;
    long rows0, rows1;
    int mode;  // 0 memcpy, 1 SSE2 stream, 2 AVX2 stream, 3 AVX-512 stream
} job_t;

__attribute__((target("avx2"))) static void row_avx2(float* d, const float* s) {
    for (int i = 0; i < ROW_FLOATS; i += 8) _mm256_stream_ps(d + i, _mm256_loadu_ps(s + i));
}
__attribute__((target("avx512f"))) static void row_avx512(float* d, const float* s) {
    for (int i = 0; i < ROW_FLOATS; i += 16) _mm512_stream_ps(d + i, _mm512_loadu_ps(s + i));
}
static void row_sse2(float* d, const float* s) {
    for (int i = 0; i < ROW_FLOATS; i += 4) _mm_stream_ps(d + i, _mm_loadu_ps(s + i));
}

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (long r = j->rows0; r < j->rows1; ++r) {
        const float* s = j->src + (size_t)r * FRAME_FLOATS;
        float* d = j->dst + (size_t)r * ROW_FLOATS;
        switch (j->mode) {
            case 0: memcpy(d, s, ROW_FLOATS * 4); break;
            case 1: row_sse2(d, s); break;
            case 2: row_avx2(d, s); break;
            default: row_avx512(d, s); break;
        }
    }
    _mm_sfence();
    return NULL;
}

static double run(const float* src, float* dst, long rows, int threads, int mode) {
    pthread_t th[256];
    job_t jobs[256];
    const double t0 = now();
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (job_t){src, dst, rows * t / threads, rows * (t + 1) / threads, mode};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    return now() - t0;
}

int main(int argc, char** argv) {
    const double gb = argc > 1 ? atof(argv[1]) : 3.0;
    const int only_threads = argc > 2 ? atoi(argv[2]) : 0;
    const long rows = (long)(gb * 1e9 / (ROW_FLOATS * 4));
    const size_t dst_bytes = (size_t)rows * ROW_FLOATS * 4, src_bytes = (size_t)(rows + 16) * FRAME_FLOATS * 4;
    const int ncpu = (int)sysconf(_SC_NPROCESSORS_ONLN);
    cpu_set_t set;
    sched_getaffinity(0, sizeof(set), &set);
    const int avail = CPU_COUNT(&set);
    const int has_avx2 = __builtin_cpu_supports("avx2"), has_512 = __builtin_cpu_supports("avx512f");
    printf("cpus online %d, usable %d, avx2 %d, avx512f %d, dst %.2f GB (%ld rows), src %.2f GB\n", ncpu, avail,
           has_avx2, has_512, dst_bytes / 1e9, rows, src_bytes / 1e9);
    float* src = (float*)malloc(src_bytes);
    for (size_t i = 0; i < src_bytes / 4; ++i) src[i] = (float)i;
    const char* names[] = {"memcpy", "sse2-nt", "avx2-nt", "avx512-nt"};
    int counts[] = {1, 2, 4, 8, 12, 16, 24, 32, 48, 64};
    // cold destination: fresh anonymous mapping per run, with and without huge pages
    for (int huge = 0; huge < 2; ++huge) {
        for (int ci = 0; ci < 10; ++ci) {
            const int th = counts[ci];
            if (th > avail || (only_threads && th != only_threads) || (th != 1 && th != avail && th != 8 && th != 16)) continue;
            float* dst = (float*)mmap(NULL, dst_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (dst == MAP_FAILED) { perror("mmap"); return 1; }
            if (huge) madvise(dst, dst_bytes, MADV_HUGEPAGE);
            const double t = run(src, dst, rows, th, has_avx2 ? 2 : 1);
            printf("cold %-9s threads %2d  %s  %7.1f ms  %6.1f GB/s\n", huge ? "hugepage" : "4k-pages", th,
                   has_avx2 ? names[2] : names[1], t * 1e3, dst_bytes / t / 1e9);
            munmap(dst, dst_bytes);
        }
    }
    float* dst = (float*)mmap(NULL, dst_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    madvise(dst, dst_bytes, MADV_HUGEPAGE);
    memset(dst, 0, dst_bytes);
    for (int mode = 0; mode < 4; ++mode) {
        if ((mode == 2 && !has_avx2) || (mode == 3 && !has_512)) continue;
        for (int ci = 0; ci < 10; ++ci) {
            const int th = counts[ci];
            if (th > avail || (only_threads && th != only_threads)) continue;
            double best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                const double t = run(src, dst, rows, th, mode);
                if (t < best) best = t;
            }
            printf("warm %-9s threads %2d  %7.1f ms  %6.1f GB/s\n", names[mode], th, best * 1e3, dst_bytes / best / 1e9);
            fflush(stdout);
        }
    }
    return 0;
}
