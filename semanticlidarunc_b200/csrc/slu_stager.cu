// slu_stager.cu -- on-disk scans to HBM: KITTI-style .bin / .label files read by native I/O threads straight into
// pinned host buffers and copied to the device on the caller's stream.
//
// Replaces (reference file:line): the two np.fromfile calls at the top of every loader's __getitem__
// (src/dataset/dataloader_semantic_KITTI.py:35-39, ..._CUDAL.py:78-84, ..._WADS.py:104-110) and the implicit
// pageable-memory H2D copy that follows when the batch is moved to the GPU (src/models/trainer.py:520-527).
// The reference hides file latency behind DataLoader worker PROCESSES that also run the numpy projection; here
// the projection runs on the GPU, so all that is left on the host is read(2) into pinned memory, done by a few
// threads that run ahead of the consumer:
//
//   submit(bin, label) -> ticket      queue a scan; an I/O thread reads both files into a free pinned slot
//   fetch(ticket, d_xyzi, d_label)    wait for that read, enqueue the two H2D copies on `stream`, record an event;
//                                     the slot is reused once the event has completed
//
// Ownership: the stager owns its pinned slots and threads (created / destroyed explicitly); device buffers belong to
// the caller, as everywhere in this ABI.  Tickets must be fetched in submission order (a fetch of the oldest
// outstanding ticket always makes progress: slots are handed to jobs strictly in ticket order).
#include <errno.h>
#include <stdlib.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "slu_common.cuh"

namespace slu {

enum SlotState { SLOT_FREE = 0, SLOT_LOADING, SLOT_READY, SLOT_COPYING };

struct StagerSlot {
    float* xyzi = nullptr;          // pinned, max_points * 4 floats
    uint32_t* label = nullptr;      // pinned, max_points
    int64_t n = 0;
    bool has_label = false;
    int state = SLOT_FREE;
    cudaEvent_t copied = nullptr;
    std::string err;
};

struct StagerJob {
    int64_t ticket;
    std::string bin, label;
};

struct Stager {
    int device = 0;
    bool host_dest = false;         // SLU_STAGER_HOST_DEST: pageable slots, destinations are host memory (I/O-logic tests)
    int64_t max_points = 0;
    std::vector<StagerSlot> slots;
    std::deque<StagerJob> jobs;
    std::map<int64_t, int> ready;   // ticket -> slot
    std::mutex mu;
    std::condition_variable cv_job, cv_ready, cv_free;
    std::vector<std::thread> threads;
    int64_t next_ticket = 0;
    bool stop = false;
};

// read exactly `bytes` from fd into dst (pinned memory: the kernel copies straight into it)
static bool read_all(int fd, void* dst, int64_t bytes, std::string& err) {
    char* p = static_cast<char*>(dst);
    int64_t done = 0;
    while (done < bytes) {
        const ssize_t r = ::read(fd, p + done, (size_t)(bytes - done));
        if (r < 0) {
            if (errno == EINTR) continue;
            err = std::string("read failed: ") + strerror(errno);
            return false;
        }
        if (r == 0) { err = "file shrank while reading"; return false; }
        done += r;
    }
    return true;
}

static void load_into(Stager& s, StagerSlot& slot, const StagerJob& job) {
    slot.err.clear();
    slot.n = 0;
    slot.has_label = false;
    int fd = ::open(job.bin.c_str(), O_RDONLY);
    if (fd < 0) { slot.err = job.bin + ": " + strerror(errno); return; }
    struct stat st;
    if (fstat(fd, &st) != 0) { slot.err = job.bin + ": " + strerror(errno); ::close(fd); return; }
    if (st.st_size % 16 != 0) {                              // np.fromfile(...).reshape(-1, 4) raises on this
        slot.err = job.bin + ": size is not a multiple of 16 bytes (float32 x,y,z,intensity records)";
        ::close(fd); return;
    }
    const int64_t n = st.st_size / 16;
    if (n > s.max_points) {
        slot.err = job.bin + ": " + std::to_string((long long)n) + " points exceed the stager's capacity of " + std::to_string((long long)s.max_points);
        ::close(fd); return;
    }
#ifdef POSIX_FADV_SEQUENTIAL
    posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
    std::string err;
    const bool ok = read_all(fd, slot.xyzi, n * 16, err);
    ::close(fd);
    if (!ok) { slot.err = job.bin + ": " + err; return; }
    if (!job.label.empty()) {
        fd = ::open(job.label.c_str(), O_RDONLY);
        if (fd < 0) { slot.err = job.label + ": " + strerror(errno); return; }
        if (fstat(fd, &st) != 0) { slot.err = job.label + ": " + strerror(errno); ::close(fd); return; }
        if (st.st_size != n * 4) {
            slot.err = job.label + ": holds " + std::to_string((long long)(st.st_size / 4)) + " labels for " + std::to_string((long long)n) + " points";
            ::close(fd); return;
        }
        const bool ok2 = read_all(fd, slot.label, n * 4, err);
        ::close(fd);
        if (!ok2) { slot.err = job.label + ": " + err; return; }
        slot.has_label = true;
    }
    slot.n = n;
}

static int find_free_slot(Stager* s) {          // caller holds s->mu
    for (size_t i = 0; i < s->slots.size(); ++i) {
        StagerSlot& sl = s->slots[i];
        if (sl.state == SLOT_COPYING && (s->host_dest || cudaEventQuery(sl.copied) == cudaSuccess)) sl.state = SLOT_FREE;
        if (sl.state == SLOT_FREE) return (int)i;
    }
    return -1;
}

static void stager_worker(Stager* s) {
    if (!s->host_dest) cudaSetDevice(s->device);
    for (;;) {
        StagerJob job;
        int si = -1;
        {
            // A job is taken from the queue only together with a slot, under one lock: slots are therefore handed out
            // in ticket order, and the oldest outstanding ticket (the one the consumer fetches next) always owns one.
            std::unique_lock<std::mutex> lk(s->mu);
            for (;;) {
                if (s->stop) return;
                if (!s->jobs.empty()) {
                    si = find_free_slot(s);
                    if (si >= 0) break;
                    s->cv_free.wait_for(lk, std::chrono::microseconds(200));     // poll the pending H2D copies
                } else {
                    s->cv_job.wait(lk);
                }
            }
            job = s->jobs.front();
            s->jobs.pop_front();
            s->slots[si].state = SLOT_LOADING;
        }
        load_into(*s, s->slots[si], job);
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->slots[si].state = SLOT_READY;
            s->ready[job.ticket] = si;
        }
        s->cv_ready.notify_all();
    }
}

}  // namespace slu

extern "C" int slu_stager_create(int n_slots, int64_t max_points_per_scan, int n_io_threads, int flags, void** out_handle) {
    using namespace slu;
    if (!out_handle) return fail(SLU_E_ARG, "out_handle is NULL");
    *out_handle = nullptr;
    if (n_slots < 1 || n_slots > 1024 || n_io_threads < 1 || n_io_threads > 64) return fail(SLU_E_RANGE, "n_slots=%d / n_io_threads=%d unsupported", n_slots, n_io_threads);
    if (max_points_per_scan < 1 || max_points_per_scan > (1LL << 28)) return fail(SLU_E_RANGE, "max_points_per_scan=%lld unsupported", (long long)max_points_per_scan);
    const bool host_dest = (flags & SLU_STAGER_HOST_DEST) != 0;
    int dev = 0;
    if (!host_dest) {
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(SLU_E_DEVICE, "no CUDA device (%s)", cudaGetErrorString(e)); }
    }
    Stager* s = new Stager();
    s->device = dev;
    s->host_dest = host_dest;
    s->max_points = max_points_per_scan;
    s->slots.resize(n_slots);
    for (auto& sl : s->slots) {
        if (host_dest) {
            sl.xyzi = static_cast<float*>(malloc((size_t)max_points_per_scan * 16));
            sl.label = static_cast<uint32_t*>(malloc((size_t)max_points_per_scan * 4));
            if (!sl.xyzi || !sl.label) {
                for (auto& t : s->slots) { free(t.xyzi); free(t.label); }
                delete s;
                return fail(SLU_E_RANGE, "slu_stager_create: out of host memory");
            }
            continue;
        }
        cudaError_t e1 = cudaHostAlloc((void**)&sl.xyzi, (size_t)max_points_per_scan * 16, cudaHostAllocDefault);
        cudaError_t e2 = e1 == cudaSuccess ? cudaHostAlloc((void**)&sl.label, (size_t)max_points_per_scan * 4, cudaHostAllocDefault) : e1;
        cudaError_t e3 = e2 == cudaSuccess ? cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming) : e2;
        if (e3 != cudaSuccess) {
            for (auto& t : s->slots) {
                if (t.xyzi) cudaFreeHost(t.xyzi);
                if (t.label) cudaFreeHost(t.label);
                if (t.copied) cudaEventDestroy(t.copied);
            }
            delete s;
            return cuda_fail(e3, "slu_stager_create: pinned allocation");
        }
    }
    for (int i = 0; i < n_io_threads; ++i) s->threads.emplace_back(stager_worker, s);
    *out_handle = s;
    return 0;
}

extern "C" int slu_stager_submit(void* handle, const char* bin_path, const char* label_path, int64_t* out_ticket) {
    using namespace slu;
    if (!handle || !bin_path || !out_ticket) return fail(SLU_E_ARG, "slu_stager_submit: NULL argument");
    Stager* s = static_cast<Stager*>(handle);
    {
        std::lock_guard<std::mutex> lk(s->mu);
        const int64_t t = s->next_ticket++;
        s->jobs.push_back(StagerJob{t, bin_path, label_path ? label_path : ""});
        *out_ticket = t;
    }
    s->cv_job.notify_one();
    return 0;
}

extern "C" int slu_stager_fetch(void* handle, int64_t ticket, float* d_xyzi, uint32_t* d_label, int64_t capacity_points,
                                int64_t* out_n_points, int* out_has_label, slu_stream_t stream) {
    using namespace slu;
    if (!handle || !d_xyzi || !out_n_points) return fail(SLU_E_ARG, "slu_stager_fetch: NULL argument");
    Stager* s = static_cast<Stager*>(handle);
    int si;
    {
        std::unique_lock<std::mutex> lk(s->mu);
        if (ticket < 0 || ticket >= s->next_ticket) return fail(SLU_E_ARG, "ticket %lld was never issued", (long long)ticket);
        s->cv_ready.wait(lk, [&] { return s->ready.count(ticket) != 0; });
        si = s->ready[ticket];
        s->ready.erase(ticket);
    }
    StagerSlot& sl = s->slots[si];
    auto release = [&](int state) {
        { std::lock_guard<std::mutex> lk(s->mu); sl.state = state; }
        s->cv_free.notify_all();
    };
    if (!sl.err.empty()) {
        const int rc = fail(SLU_E_IO, "%s", sl.err.c_str());
        release(SLOT_FREE);
        return rc;
    }
    if (sl.n > capacity_points) {
        const int rc = fail(SLU_E_RANGE, "scan has %lld points, destination holds %lld", (long long)sl.n, (long long)capacity_points);
        release(SLOT_FREE);
        return rc;
    }
    if (s->host_dest) {                                       // destinations are host pointers: plain copies, slot free at once
        if (sl.n > 0) {
            memcpy(d_xyzi, sl.xyzi, (size_t)sl.n * 16);
            if (sl.has_label && d_label) memcpy(d_label, sl.label, (size_t)sl.n * 4);
        }
        *out_n_points = sl.n;
        if (out_has_label) *out_has_label = sl.has_label ? 1 : 0;
        release(SLOT_FREE);
        return 0;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaSuccess;
    if (sl.n > 0) {
        e = cudaMemcpyAsync(d_xyzi, sl.xyzi, (size_t)sl.n * 16, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && sl.has_label && d_label) e = cudaMemcpyAsync(d_label, sl.label, (size_t)sl.n * 4, cudaMemcpyHostToDevice, st);
    }
    if (e == cudaSuccess) e = cudaEventRecord(sl.copied, st);
    *out_n_points = sl.n;
    if (out_has_label) *out_has_label = sl.has_label ? 1 : 0;
    if (e != cudaSuccess) { release(SLOT_FREE); return cuda_fail(e, "slu_stager_fetch: H2D copy"); }
    release(SLOT_COPYING);
    return 0;
}

extern "C" int slu_stager_destroy(void* handle) {
    using namespace slu;
    if (!handle) return 0;
    Stager* s = static_cast<Stager*>(handle);
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->stop = true;
    }
    s->cv_job.notify_all();
    s->cv_free.notify_all();
    for (auto& t : s->threads) t.join();
    for (auto& sl : s->slots) {
        if (s->host_dest) { free(sl.xyzi); free(sl.label); continue; }
        if (sl.copied) { cudaEventSynchronize(sl.copied); cudaEventDestroy(sl.copied); }
        if (sl.xyzi) cudaFreeHost(sl.xyzi);
        if (sl.label) cudaFreeHost(sl.label);
    }
    delete s;
    return 0;
}
