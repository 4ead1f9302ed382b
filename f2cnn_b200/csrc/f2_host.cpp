// f2_host.cpp -- the HOST half of the window stage (include/f2cnn_b200.h, "host side").
//
// Row k of an utterance in the reference's input tensor (scripts/processing/InputGenerator.py:
// 73-80) is the envelope at the 2R+1 samples center + STEP*(j - R): when the centres lie on the
// decimated grid those are 2R+1 CONSECUTIVE decimated frames, i.e. one contiguous block of
// (2R+1)*C floats of the [frame][C] matrix the fused kernel produces.  The (N, 2R+1, C) tensor is
// therefore (2R+1)x redundant, and shipping it over PCIe costs 11x the bytes of the frames
// themselves.  This file holds what lets the frames travel instead:
//   f2_window_runs    label timepoints -> runs of consecutive windows (index checks included)
//   f2_place_windows  frames -> rows: one contiguous copy per row, non-temporal stores, threads
//   f2_placer_*       the same as an asynchronous worker pool fed in CUDA stream order
//   f2_host_alloc     huge-page backed host memory for a fresh output tensor
//   f2_host_pin       ... page-locked and kernel-addressable (the frames buffer of the corpus pipeline)
// No arithmetic happens here: placement copies float32 bit patterns the device computed.
#include <cuda_runtime.h>
#include <immintrin.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/f2cnn_b200.h"

extern "C" int f2_set_error(int code, const char* fmt, ...);  // f2_capi.cu: thread-local message

namespace {

// ---- one row: dst is written once and never read back here -> non-temporal stores ------------
__attribute__((target("avx2"))) void copy_row_avx2(float* dst, const float* src, size_t bytes) {
    char* d = reinterpret_cast<char*>(dst);
    const char* s = reinterpret_cast<const char*>(src);
    const size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
    if (head) {
        const size_t h = std::min(head, bytes);
        memcpy(d, s, h);
        d += h, s += h, bytes -= h;
    }
    for (; bytes >= 128; bytes -= 128, d += 128, s += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 96), e);
    }
    for (; bytes >= 32; bytes -= 32, d += 32, s += 32)
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d), _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s)));
    if (bytes) memcpy(d, s, bytes);
}

void copy_row_sse2(float* dst, const float* src, size_t bytes) {
    char* d = reinterpret_cast<char*>(dst);
    const char* s = reinterpret_cast<const char*>(src);
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
    if (head) {
        const size_t h = std::min(head, bytes);
        memcpy(d, s, h);
        d += h, s += h, bytes -= h;
    }
    for (; bytes >= 16; bytes -= 16, d += 16, s += 16)
        _mm_stream_si128(reinterpret_cast<__m128i*>(d), _mm_loadu_si128(reinterpret_cast<const __m128i*>(s)));
    if (bytes) memcpy(d, s, bytes);
}

typedef void (*copy_fn)(float*, const float*, size_t);
copy_fn pick_copy() {
    static copy_fn fn = __builtin_cpu_supports("avx2") ? copy_row_avx2 : copy_row_sse2;
    return fn;
}

// ---- a placement job and its decomposition into chunks of rows ----------------------------------
struct Job {
    const float* frames = nullptr;
    float* out = nullptr;
    int C = 0, dots = 0;
    std::vector<f2_win_run> runs;
    std::vector<int64_t> cum;  // rows before run i (local numbering), size runs+1
    int64_t n_chunks = 0;
    std::atomic<int64_t> next{0};
    std::atomic<int64_t> done{0};
    double t_submit = 0.0, t_ready = 0.0;  // trace: monotonic seconds
};

double now_s() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

constexpr int64_t kChunkRows = 192;  // ~1 MB of output per chunk at 11 x 128 floats per row

void prepare(Job& j) {
    j.cum.resize(j.runs.size() + 1);
    j.cum[0] = 0;
    for (size_t i = 0; i < j.runs.size(); ++i) j.cum[i + 1] = j.cum[i] + j.runs[i].count;
    j.n_chunks = (j.cum.back() + kChunkRows - 1) / kChunkRows;
}

// rows [r0, r1) of the job's local numbering
void place_rows(const Job& j, int64_t r0, int64_t r1) {
    const copy_fn copy = pick_copy();
    const size_t row_floats = (size_t)j.dots * (size_t)j.C;
    size_t i = (size_t)(std::upper_bound(j.cum.begin(), j.cum.end(), r0) - j.cum.begin()) - 1;
    while (r0 < r1) {
        const f2_win_run& run = j.runs[i];
        const int64_t k0 = r0 - j.cum[i];
        const int64_t k1 = std::min<int64_t>(run.count, r1 - j.cum[i]);
        const float* src = j.frames + (size_t)(run.first_frame + k0) * (size_t)j.C;
        float* dst = j.out + (size_t)(run.row0 + k0) * row_floats;
        for (int64_t k = k0; k < k1; ++k, src += j.C, dst += row_floats) copy(dst, src, row_floats * sizeof(float));
        r0 = j.cum[i] + k1;
        ++i;
    }
}

bool work_on(Job& j) {  // returns true when this call finished the job's last chunk
    bool last = false;
    for (;;) {
        const int64_t c = j.next.fetch_add(1, std::memory_order_relaxed);
        if (c >= j.n_chunks) break;
        const int64_t r0 = c * kChunkRows;
        place_rows(j, r0, std::min(j.cum.back(), r0 + kChunkRows));
        _mm_sfence();
        if (j.done.fetch_add(1, std::memory_order_acq_rel) + 1 == j.n_chunks) last = true;
    }
    return last;
}

int validate(const float* frames, int C, int dots, const f2_win_run* runs, int64_t n_runs, const float* out) {
    if (n_runs < 0 || C <= 0 || dots <= 0 || (n_runs > 0 && (!frames || !runs || !out)))
        return f2_set_error(F2_ERR_INVALID, "window placement: bad arguments");
    for (int64_t i = 0; i < n_runs; ++i)
        if (runs[i].count < 0 || runs[i].first_frame < 0 || runs[i].row0 < 0)
            return f2_set_error(F2_ERR_INVALID, "window placement: run %lld is negative", (long long)i);
    return F2_OK;
}

int default_threads() {
    cpu_set_t set;
    int n = 0;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    return std::max(1, n);
}

}  // namespace

// ---- the asynchronous pool ---------------------------------------------------------------------
struct f2_placer {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_idle;
    std::deque<std::shared_ptr<Job>> ready;  // runnable jobs (workers share the front one)
    int64_t pending = 0;                     // submitted and not yet finished
    bool stop = false;
    std::vector<double> trace;               // per finished job: submit, ready, done (seconds), rows

    void loop() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || !ready.empty(); });
            if (ready.empty()) return;  // stop
            std::shared_ptr<Job> j = ready.front();
            if (j->next.load(std::memory_order_relaxed) >= j->n_chunks) {
                // every chunk is taken: whoever finishes the last one retires the job
                ready.pop_front();
                continue;
            }
            lk.unlock();
            const bool last = work_on(*j);
            lk.lock();
            if (last) {
                if (!ready.empty() && ready.front() == j) ready.pop_front();
                if (trace.size() < 4 * 4096) {
                    const double t[4] = {j->t_submit, j->t_ready, now_s(), (double)j->cum.back()};
                    trace.insert(trace.end(), t, t + 4);
                }
                if (--pending == 0) cv_idle.notify_all();
            }
        }
    }
    void make_ready(std::shared_ptr<Job> j) {
        std::lock_guard<std::mutex> lk(mu);
        j->t_ready = now_s();
        if (j->n_chunks == 0) {
            if (--pending == 0) cv_idle.notify_all();
            return;
        }
        ready.push_back(std::move(j));
        cv_work.notify_all();
    }
};

namespace {
struct HostFuncArg {
    f2_placer* placer;
    std::shared_ptr<Job> job;
};
void CUDART_CB on_stream_reached(void* p) {  // runs on a CUDA driver thread: no CUDA calls in here
    HostFuncArg* a = static_cast<HostFuncArg*>(p);
    a->placer->make_ready(std::move(a->job));
    delete a;
}
}  // namespace

extern "C" {

int f2_window_runs(const int64_t* centers, const int64_t* counts, const int64_t* lengths, const int64_t* frame_offsets,
                   const int64_t* row_offsets, int n_utts, int radius, int step, int* phase, f2_win_run* runs,
                   int64_t max_runs, int64_t* n_runs, int64_t* n_rows) {
    if (n_utts < 0 || radius < 0 || step <= 0 || !phase || !n_runs || !n_rows ||
        (n_utts > 0 && (!counts || !lengths || !frame_offsets)))
        return f2_set_error(F2_ERR_INVALID, "f2_window_runs: bad arguments");
    const int64_t reach = (int64_t)radius * step;
    int64_t ph = *phase;  // < 0: take it from the first window
    int64_t nr = 0, rows = 0, pos = 0;
    bool strided = true;
    for (int u = 0; u < n_utts; ++u) {
        const int64_t cnt = counts[u], n = lengths[u];
        if (cnt < 0 || n < 0 || (cnt > 0 && !centers))
            return f2_set_error(F2_ERR_INVALID, "f2_window_runs: utterance %d has a negative count or length", u);
        int64_t row = row_offsets ? row_offsets[u] : rows;
        int64_t prev_first = INT64_MIN / 2;  // first sample of the previous window IF it extended/opened a run
        const int64_t last_ok = n - 1 - reach;  // largest centre whose last sample is inside the utterance
        const int64_t* cen = centers + pos;
        for (int64_t i = 0; i < cnt; ++i, ++row) {
            const int64_t c = cen[i];
            const int64_t first = c - reach;
            if (first == prev_first + step && c <= last_ok) {
                // the common case (a label grid): next frame of the same run -- no division, no checks
                if (runs && nr <= max_runs) runs[nr - 1].count += 1;
                prev_first = first;
                continue;
            }
            prev_first = INT64_MIN / 2;
            // InputGenerator.py:76 indexes Python lists/arrays: an index in [-n, 0) wraps, anything else
            // outside [0, n) is an IndexError
            if (c > last_ok || first < -n)
                return f2_set_error(F2_ERR_INDEX, "index out of bounds: utterance %d, timepoint %lld, %lld samples", u,
                                    (long long)c, (long long)n);
            if (first < 0) {  // wraps around the end: not a block of consecutive frames
                strided = false;
                continue;
            }
            if (ph < 0) ph = first % step;
            if (first % step != ph) {
                strided = false;
                continue;
            }
            if (!strided) continue;
            if (runs && nr < max_runs) {
                runs[nr].first_frame = frame_offsets[u] + (first - ph) / step;
                runs[nr].row0 = row;
                runs[nr].count = 1;
            }
            ++nr;
            prev_first = first;
        }
        pos += cnt;
        rows += cnt;
    }
    *n_rows = rows;
    *phase = (int)(ph < 0 ? 0 : ph);
    if (!strided) {
        *n_runs = -1;  // legal timepoints, but not windows of consecutive frames of one grid
        return F2_OK;
    }
    *n_runs = nr;
    if (runs && nr > max_runs) return f2_set_error(F2_ERR_WORKSPACE, "f2_window_runs: %lld runs, room for %lld", (long long)nr, (long long)max_runs);
    return F2_OK;
}

int f2_place_windows(const float* frames, int n_channels, int dots, const f2_win_run* runs, int64_t n_runs, float* out,
                     int n_threads) {
    const int rc = validate(frames, n_channels, dots, runs, n_runs, out);
    if (rc != F2_OK || n_runs == 0) return rc;
    Job j;
    j.frames = frames, j.out = out, j.C = n_channels, j.dots = dots;
    j.runs.assign(runs, runs + n_runs);
    prepare(j);
    if (n_threads <= 0) n_threads = default_threads();
    n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, j.n_chunks));
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back([&j] { work_on(j); });
    work_on(j);
    for (auto& t : th) t.join();
    return F2_OK;
}

int f2_placer_create(int n_threads, f2_placer** out) {
    if (!out) return f2_set_error(F2_ERR_INVALID, "f2_placer_create: null output");
    if (n_threads <= 0) n_threads = default_threads();
    f2_placer* p = new (std::nothrow) f2_placer();
    if (!p) return f2_set_error(F2_ERR_INVALID, "out of host memory");
    try {
        for (int t = 0; t < n_threads; ++t) p->workers.emplace_back([p] { p->loop(); });
    } catch (...) {
        {
            std::lock_guard<std::mutex> lk(p->mu);
            p->stop = true;
        }
        p->cv_work.notify_all();
        for (auto& t : p->workers) t.join();
        delete p;
        return f2_set_error(F2_ERR_INVALID, "f2_placer_create: cannot start %d threads", n_threads);
    }
    *out = p;
    return F2_OK;
}

int f2_placer_threads(const f2_placer* p) { return p ? (int)p->workers.size() : 0; }

int f2_placer_submit(f2_placer* p, int after_stream, void* stream, const float* frames, int n_channels, int dots,
                     const f2_win_run* runs, int64_t n_runs, float* out) {
    if (!p) return f2_set_error(F2_ERR_INVALID, "f2_placer_submit: null placer");
    const int rc = validate(frames, n_channels, dots, runs, n_runs, out);
    if (rc != F2_OK || n_runs == 0) return rc;
    std::shared_ptr<Job> j = std::make_shared<Job>();
    j->frames = frames, j->out = out, j->C = n_channels, j->dots = dots;
    j->runs.assign(runs, runs + n_runs);
    prepare(*j);
    j->t_submit = now_s();
    {
        std::lock_guard<std::mutex> lk(p->mu);
        ++p->pending;
    }
    if (!after_stream) {
        p->make_ready(std::move(j));
        return F2_OK;
    }
    HostFuncArg* a = new (std::nothrow) HostFuncArg{p, j};
    cudaError_t e = a ? cudaLaunchHostFunc((cudaStream_t)stream, on_stream_reached, a) : cudaErrorMemoryAllocation;
    if (e != cudaSuccess) {
        delete a;
        std::lock_guard<std::mutex> lk(p->mu);
        if (--p->pending == 0) p->cv_idle.notify_all();
        return f2_set_error(F2_ERR_CUDA, "cudaLaunchHostFunc: %s", cudaGetErrorString(e));
    }
    return F2_OK;
}

int f2_placer_wait(f2_placer* p) {
    if (!p) return f2_set_error(F2_ERR_INVALID, "f2_placer_wait: null placer");
    std::unique_lock<std::mutex> lk(p->mu);
    p->cv_idle.wait(lk, [&] { return p->pending == 0; });
    return F2_OK;
}

int64_t f2_placer_trace(f2_placer* p, double* out, int64_t max_jobs, int clear) {
    if (!p) return 0;
    std::lock_guard<std::mutex> lk(p->mu);
    const int64_t n = std::min<int64_t>((int64_t)p->trace.size() / 4, out ? max_jobs : 0);
    for (int64_t i = 0; i < 4 * n; ++i) out[i] = p->trace[(size_t)i];
    const int64_t have = (int64_t)p->trace.size() / 4;
    if (clear) p->trace.clear();
    return out ? n : have;
}

int f2_placer_destroy(f2_placer* p) {
    if (!p) return F2_OK;
    f2_placer_wait(p);
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv_work.notify_all();
    for (auto& t : p->workers) t.join();
    delete p;
    return F2_OK;
}

// ---- host memory for a fresh output tensor ------------------------------------------------------
// A never-touched malloc'ed tensor of 7.5 GB costs ~2 M page faults on its first write; an anonymous
// mapping advised to use transparent huge pages takes 512x fewer.
int f2_host_alloc(size_t bytes, void** out) {
    if (!out) return f2_set_error(F2_ERR_INVALID, "f2_host_alloc: null output");
    const size_t len = std::max<size_t>(bytes, 1);
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return f2_set_error(F2_ERR_INVALID, "f2_host_alloc: mmap of %zu bytes failed", len);
#ifdef MADV_HUGEPAGE
    if (len >= ((size_t)2 << 20)) madvise(p, len, MADV_HUGEPAGE);
#endif
    *out = p;
    return F2_OK;
}

// Page-lock a mapping from f2_host_alloc for the current device and make it addressable from kernels under
// the SAME pointer: pages are faulted in by a few threads first (huge pages: 0.7 GB in tens of ms), then
// registered -- a quarter of the time cudaHostAlloc takes for the same bytes (tools/pin_probe.py).
int f2_host_pin(void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return f2_set_error(F2_ERR_INVALID, "f2_host_pin: bad arguments");
    {
        const int nt = 4;
        std::vector<std::thread> th;
        const size_t per = (bytes + nt - 1) / nt;
        for (int i = 0; i < nt; ++i)
            th.emplace_back([=] {
                volatile char* c = (volatile char*)ptr;
                const size_t end = std::min(bytes, (size_t)(i + 1) * per);
                for (size_t o = (size_t)i * per; o < end; o += 4096) c[o] = 0;
            });
        for (auto& t : th) t.join();
    }
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return f2_set_error(F2_ERR_CUDA, "f2_host_pin: cudaHostRegister of %zu bytes: %s", bytes, cudaGetErrorString(e));
    }
    void* dptr = nullptr;
    e = cudaHostGetDevicePointer(&dptr, ptr, 0);
    if (e != cudaSuccess || dptr != ptr) {
        cudaGetLastError();
        cudaHostUnregister(ptr);
        return f2_set_error(F2_ERR_UNSUPPORTED, "f2_host_pin: registered host memory is not addressable from the device "
                                                "under its host pointer");
    }
    return F2_OK;
}

int f2_host_unpin(void* ptr) {
    if (ptr && cudaHostUnregister(ptr) != cudaSuccess) {
        cudaGetLastError();
        return f2_set_error(F2_ERR_CUDA, "f2_host_unpin: cudaHostUnregister failed");
    }
    return F2_OK;
}

int f2_host_free(void* ptr, size_t bytes) {
    if (ptr && munmap(ptr, std::max<size_t>(bytes, 1)) != 0) return f2_set_error(F2_ERR_INVALID, "f2_host_free: munmap failed");
    return F2_OK;
}

}  // extern "C"
