// host_pipeline.h -- the host side of a device's data path, shared by the fused-chain plan (mmsgpu_api.cu) and the
// stage-level twins (api_twins.cu).
//
// A Pipe is bound to one device.  Its work is cut into chunks (column windows of [...][cell] arrays for the chain,
// blocks of whole pixels for the twins) and a small team of host threads ("workers") moves them:
//
//      worker w:   stage + H2D chunk k      | drain chunk k-1        | stage + H2D chunk k+1 | ...
//      device  :   kernels of chunk k-1     | kernels of chunk k     | kernels of chunk k+1  | ...
//
// Every worker moves its share of the columns of every array (Mat) of the chunk.  The last worker to deliver its
// part of chunk k enqueues that chunk's kernels (a callback of the job) on the pipe's compute stream behind the
// workers' copy events; each worker's download of its columns waits on the chunk's kernel event.  Host memory the
// driver can DMA directly (mms_host_alloc, cudaHostRegister) is copied in place with one cudaMemcpy2DAsync per array
// and chunk; ordinary pageable memory -- what R vectors and numpy arrays are -- goes through per-worker rings of
// pinned bounce buffers (a worker's memcpy of piece j overlaps the DMA of piece j +- 1; streaming stores,
// host_copy.cpp), several workers per device, because one thread's memcpy is slower than the PCIe link.
#pragma once

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

#include "api_common.h"

namespace mms_api {

constexpr size_t kSlotBytes = (size_t)16 << 20;   // one pinned bounce buffer
constexpr int kSlots = 2;                         // bounce buffers per worker and direction
constexpr int kMaxWorkers = 12;                   // host threads per device

// host_copy.cpp: streaming-store row copy between caller memory and the pinned bounce buffers
void copy_rows(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, size_t rows);
// columns [w0, w0 + wn) of a chunk (first column c0, n_cells columns) that worker w of `workers` moves: whole units of
// `unit` columns except at the chunk's end; the windows of all workers tile the chunk (host_copy.cpp)
void window_of_worker(int64_t c0, int64_t n_cells, int w, int workers, int64_t unit, int64_t &w0, int64_t &wn);

// Process-wide pool of the pinned bounce buffers (kSlotBytes each): page-locking memory is slow (milliseconds per
// buffer), and the stage-level twins build a pipe per call.  Buffers go back to the pool when a pipe stops and are
// kept (at most kPoolKeep of them) for the next one.
void *pinned_slot_acquire();      // nullptr on failure (the CUDA error is left for cudaGetLastError)
void pinned_slot_release(void *p);

// a host <-> device array whose fastest index is the "cell" (the unit the chunks cut): `rows` rows of cells
struct Mat {
    char *host = nullptr;       // the pipe's first cell of row 0 on the host
    size_t host_pitch = 0;      // bytes between rows on the host
    char *dev = nullptr;        // the same on the device
    size_t dev_pitch = 0;
    size_t rows = 0;
    size_t elem = 8;            // bytes per cell
    bool pinned = false;        // the driver can DMA the host side directly
};

struct PipeChunk {
    int64_t c0 = 0;       // first cell (column) within the pipe's arrays
    int64_t n_cols = 0;   // columns the chunk's kernels cover (padded)
    int64_t n_cells = 0;  // real cells among them (what is copied)
    cudaEvent_t ev_kernel = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    std::vector<cudaEvent_t> ev_in;   // [workers]
    bool kernel_recorded = false;
};

enum PipeKind { PIPE_NONE = 0, PIPE_CALL, PIPE_UPLOAD, PIPE_RUN, PIPE_DOWNLOAD, PIPE_EXIT };

struct PipeJob {
    PipeKind kind = PIPE_NONE;
    std::vector<Mat> in, out;
    int workers = 1;                                     // workers that move data (1 when everything is pinned)
    std::function<bool()> call;                          // PIPE_CALL: runs on worker 0 with the device current
    std::function<bool(cudaStream_t)> before_first_copy; // worker 0, on its copy stream, before chunk 0
    std::function<bool(int)> inputs_ready;               // the chunk's inputs are on their way (the compute stream waits
                                                         // for them): enqueue what derives the kernel's inputs from them
    std::function<bool(int)> run_chunk;                  // PIPE_RUN: enqueue the chunk's kernels (the pipe brackets them
                                                         // with the chunk's timing events and records its kernel event)
    // output columns [o0, o0 + on) that go with the cell window [w0, w0 + wn) (default: the same window)
    std::function<void(int64_t, int64_t, int64_t &, int64_t &)> out_window;
};

class Pipe {
public:
    int device = 0;
    cudaStream_t s_compute = nullptr;
    std::vector<PipeChunk> chunks;
    int64_t window_unit = 32;                  // workers split a chunk's cells at multiples of this many cells

    Pipe();
    ~Pipe();
    Pipe(const Pipe &) = delete;
    Pipe &operator=(const Pipe &) = delete;
    void start(int n_workers);                 // spawn the workers (each makes `device` current)
    void stop();                               // join them (device objects of the workers are released by them)
    int n_workers() const { return (int)workers_.size(); }
    bool create_device_objects();              // compute stream + chunk events; call with the device current
    void destroy_device_objects();
    void post(const PipeJob *job);             // hand the job to the workers ...
    int wait(std::string *message);            // ... and wait for all of them; MMS_OK or the first failure
    int run(const PipeJob *job, std::string *message) { post(job); return wait(message); }
    void set_error(int status, const std::string &msg);
    bool failed() const { return failed_.load(); }
    void clear_error();

private:
    struct Worker;
    void worker_main(Worker *wk);
    void run_job(Worker *wk, const PipeJob &job);
    int team_size(const PipeJob &job) const;
    void arrive(Worker *wk, const PipeJob &job, int k);
    bool do_in(Worker *wk, const PipeJob &job, int k);
    bool do_out(Worker *wk, const PipeJob &job, int k);
    bool stage_in(Worker *wk, const Mat &m, int64_t w0, int64_t wn);
    bool stage_out(Worker *wk, const Mat &m, int64_t w0, int64_t wn);
    bool drain_one(Worker *wk);
    bool ensure_slots(Worker *wk, bool in);

    std::vector<std::unique_ptr<Worker>> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    uint64_t job_seq_ = 0;
    const PipeJob *job_ = nullptr;
    int pending_ = 0;                               // workers still busy with the current job
    std::unique_ptr<std::atomic<int>[]> arrived_;   // [chunk] workers that delivered their inputs
    size_t arrived_size_ = 0;
    std::vector<char> enqueued_;                    // [chunk] kernels of the chunk are on the compute stream (guarded by mu_)
    std::atomic<bool> failed_{false};
    int err_status_ = MMS_OK;
    std::string err_msg_;
};

// CUDA call inside the pipeline: record the failure in the pipe and leave with false
#define MMS_CUDA_P(pipe, call)                                                                    \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            char b_[512];                                                                         \
            snprintf(b_, sizeof b_, "CUDA error %s at %s:%d (%s) on device %d", cudaGetErrorName(e_), __FILE__, \
                     __LINE__, cudaGetErrorString(e_), (pipe)->device);                           \
            (pipe)->set_error(e_ == cudaErrorMemoryAllocation ? MMS_ERR_ALLOC : MMS_ERR_CUDA, b_); \
            return false;                                                                         \
        }                                                                                         \
    } while (0)

// columns [w0, w0 + wn) of a chunk that worker `w` of `workers` moves (real cells only, whole `unit`-cell blocks)
void worker_window(const PipeChunk &c, int w, int workers, int64_t unit, int64_t &w0, int64_t &wn);

// host threads per device: MMS_HOST_WORKERS, else the hardware threads shared among the devices of the call
int default_workers(int n_devices);

// pinned or pageable?  One driver query per host array; returns true when every array is pinned
bool classify_mats(std::vector<Mat> &mats);

}  // namespace mms_api
