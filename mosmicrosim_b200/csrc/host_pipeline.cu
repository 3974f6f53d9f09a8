// host_pipeline.cu -- see host_pipeline.h.
#include "host_pipeline.h"

namespace mms_api {

namespace {
constexpr size_t kPoolKeep = 64;            // 1 GB of page-locked memory at most stays with the process
std::mutex g_pool_mu;
std::vector<void *> g_pool;
}  // namespace

void *pinned_slot_acquire()
{
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_pool.empty()) { void *p = g_pool.back(); g_pool.pop_back(); return p; }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, kSlotBytes, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    return p;
}

void pinned_slot_release(void *p)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_pool.size() < kPoolKeep) { g_pool.push_back(p); return; }
    }
    cudaFreeHost(p);
}

struct Pipe::Worker {
    int id = 0;
    std::thread th;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    void *pin_in[kSlots] = {nullptr, nullptr}, *pin_out[kSlots] = {nullptr, nullptr};
    cudaEvent_t ev_slot_in[kSlots] = {nullptr, nullptr}, ev_slot_out[kSlots] = {nullptr, nullptr};
    bool slot_in_used[kSlots] = {false, false};
    int next_in = 0, next_out = 0;
    struct Pending { int slot; char *dst; size_t pitch, width, rows; };
    std::deque<Pending> pending;
};

Pipe::Pipe() {}
Pipe::~Pipe() { stop(); }

void worker_window(const PipeChunk &c, int w, int workers, int64_t unit, int64_t &w0, int64_t &wn)
{
    window_of_worker(c.c0, c.n_cells, w, workers, unit, w0, wn);      // host_copy.cpp (CPU-tested)
}

int default_workers(int n_devices)
{
    if (const char *e = std::getenv("MMS_HOST_WORKERS")) {
        const int w = std::atoi(e);
        if (w >= 1) return std::min(w, 16);
    }
    // the hardware threads are shared among the devices of the call -- and, in a one-process-per-GPU launch
    // (torchrun / mpirun export the number of local processes), among those processes
    int procs = 1;
    for (const char *name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS"})
        if (const char *e = std::getenv(name)) { const int v = std::atoi(e); if (v > 1) { procs = v; break; } }
    const unsigned hw = std::thread::hardware_concurrency();
    const int per = hw ? (int)hw / std::max(1, n_devices * procs) : 2;
    return std::max(1, std::min(kMaxWorkers, per));
}

bool classify_mats(std::vector<Mat> &mats)
{
    bool all = true;
    for (Mat &m : mats) { m.pinned = host_pointer_is_pinned(m.host); all = all && m.pinned; }
    return all;
}

// the workers that take part in a job: the first job.workers of the started ones (at least one)
int Pipe::team_size(const PipeJob &job) const
{
    return std::max(1, std::min(job.workers, (int)workers_.size()));
}

void Pipe::set_error(int status, const std::string &msg)
{
    std::lock_guard<std::mutex> lk(mu_);
    if (!failed_.load()) { err_status_ = status; err_msg_ = msg; failed_.store(true); }
    cv_.notify_all();
}

void Pipe::clear_error()
{
    std::lock_guard<std::mutex> lk(mu_);
    failed_.store(false);
    err_status_ = MMS_OK;
    err_msg_.clear();
}

bool Pipe::create_device_objects()
{
    MMS_CUDA_P(this, cudaStreamCreateWithFlags(&s_compute, cudaStreamNonBlocking));
    for (PipeChunk &c : chunks) {
        MMS_CUDA_P(this, cudaEventCreateWithFlags(&c.ev_kernel, cudaEventDisableTiming));
        MMS_CUDA_P(this, cudaEventCreate(&c.ev_t0));
        MMS_CUDA_P(this, cudaEventCreate(&c.ev_t1));
        c.ev_in.assign(workers_.size(), nullptr);
        for (cudaEvent_t &e : c.ev_in) MMS_CUDA_P(this, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return true;
}

void Pipe::destroy_device_objects()
{
    for (PipeChunk &c : chunks) {
        if (c.ev_kernel) cudaEventDestroy(c.ev_kernel);
        if (c.ev_t0) cudaEventDestroy(c.ev_t0);
        if (c.ev_t1) cudaEventDestroy(c.ev_t1);
        for (cudaEvent_t e : c.ev_in) if (e) cudaEventDestroy(e);
        c.ev_kernel = c.ev_t0 = c.ev_t1 = nullptr;
        c.ev_in.clear();
    }
    if (s_compute) cudaStreamDestroy(s_compute);
    s_compute = nullptr;
}

void Pipe::start(int n_workers)
{
    arrived_.reset(new std::atomic<int>[std::max<size_t>(1, chunks.size())]);
    arrived_size_ = chunks.size();
    enqueued_.assign(chunks.size(), 0);
    for (int w = 0; w < n_workers; ++w) {
        std::unique_ptr<Worker> wk(new Worker());
        wk->id = w;
        workers_.push_back(std::move(wk));
    }
    for (auto &wk : workers_) wk->th = std::thread(&Pipe::worker_main, this, wk.get());
}

void Pipe::stop()
{
    if (workers_.empty()) return;
    PipeJob bye;
    bye.kind = PIPE_EXIT;
    {
        std::lock_guard<std::mutex> lk(mu_);
        job_ = &bye;
        job_seq_ += 1;
    }
    cv_.notify_all();
    for (auto &w : workers_)
        if (w->th.joinable()) w->th.join();
    workers_.clear();
}

void Pipe::post(const PipeJob *job)
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (size_t k = 0; k < chunks.size(); ++k) {
            arrived_[k].store(0);
            enqueued_[k] = 0;
            if (job->kind == PIPE_RUN || job->kind == PIPE_UPLOAD) chunks[k].kernel_recorded = false;
        }
        failed_.store(false);
        err_status_ = MMS_OK;
        err_msg_.clear();
        job_ = job;
        pending_ = (int)workers_.size();
        job_seq_ += 1;
    }
    cv_.notify_all();
}

int Pipe::wait(std::string *message)
{
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return pending_ == 0; });
    if (failed_.load()) {
        if (message) *message = err_msg_;
        return err_status_;
    }
    return MMS_OK;
}

bool Pipe::ensure_slots(Worker *wk, bool in)
{
    void **bufs = in ? wk->pin_in : wk->pin_out;
    cudaEvent_t *evs = in ? wk->ev_slot_in : wk->ev_slot_out;
    for (int k = 0; k < kSlots; ++k) {
        if (!bufs[k]) {
            bufs[k] = pinned_slot_acquire();
            if (!bufs[k]) { set_error(MMS_ERR_ALLOC, "could not page-lock a bounce buffer (cudaHostAlloc)"); cudaGetLastError(); return false; }
        }
        if (!evs[k]) MMS_CUDA_P(this, cudaEventCreateWithFlags(&evs[k], cudaEventDisableTiming));
    }
    return true;
}

// host -> device: the worker's columns of one array
bool Pipe::stage_in(Worker *wk, const Mat &m, int64_t w0, int64_t wn)
{
    if (wn <= 0 || m.rows == 0) return true;
    const size_t off = (size_t)w0 * m.elem, width = (size_t)wn * m.elem;
    if (m.pinned) {
        if (m.rows == 1)
            MMS_CUDA_P(this, cudaMemcpyAsync(m.dev + off, m.host + off, width, cudaMemcpyHostToDevice, wk->s_in));
        else
            MMS_CUDA_P(this, cudaMemcpy2DAsync(m.dev + off, m.dev_pitch, m.host + off, m.host_pitch, width, m.rows,
                                               cudaMemcpyHostToDevice, wk->s_in));
        return true;
    }
    if (!ensure_slots(wk, true)) return false;
    const size_t col_step = std::min(width, kSlotBytes);
    for (size_t col = 0; col < width; col += col_step) {
        const size_t wd = std::min(col_step, width - col);
        const size_t rows_per = std::max<size_t>(1, kSlotBytes / wd);
        for (size_t r0 = 0; r0 < m.rows; r0 += rows_per) {
            const size_t nr = std::min(rows_per, m.rows - r0);
            const int slot = wk->next_in;
            wk->next_in = (slot + 1) % kSlots;
            if (wk->slot_in_used[slot]) MMS_CUDA_P(this, cudaEventSynchronize(wk->ev_slot_in[slot]));
            char *buf = (char *)wk->pin_in[slot];
            copy_rows(buf, wd, m.host + r0 * m.host_pitch + off + col, m.host_pitch, wd, nr);
            if (nr == 1)
                MMS_CUDA_P(this, cudaMemcpyAsync(m.dev + r0 * m.dev_pitch + off + col, buf, wd, cudaMemcpyHostToDevice, wk->s_in));
            else
                MMS_CUDA_P(this, cudaMemcpy2DAsync(m.dev + r0 * m.dev_pitch + off + col, m.dev_pitch, buf, wd, wd, nr,
                                                   cudaMemcpyHostToDevice, wk->s_in));
            MMS_CUDA_P(this, cudaEventRecord(wk->ev_slot_in[slot], wk->s_in));
            wk->slot_in_used[slot] = true;
        }
    }
    return true;
}

bool Pipe::drain_one(Worker *wk)
{
    const Worker::Pending q = wk->pending.front();
    wk->pending.pop_front();
    MMS_CUDA_P(this, cudaEventSynchronize(wk->ev_slot_out[q.slot]));
    copy_rows(q.dst, q.pitch, (const char *)wk->pin_out[q.slot], q.width, q.width, q.rows);
    return true;
}

// device -> host: the worker's columns of one array (the stream already waits for the producing kernel)
bool Pipe::stage_out(Worker *wk, const Mat &m, int64_t w0, int64_t wn)
{
    if (wn <= 0 || m.rows == 0) return true;
    const size_t off = (size_t)w0 * m.elem, width = (size_t)wn * m.elem;
    if (m.pinned) {
        if (m.rows == 1)
            MMS_CUDA_P(this, cudaMemcpyAsync(m.host + off, m.dev + off, width, cudaMemcpyDeviceToHost, wk->s_out));
        else
            MMS_CUDA_P(this, cudaMemcpy2DAsync(m.host + off, m.host_pitch, m.dev + off, m.dev_pitch, width, m.rows,
                                               cudaMemcpyDeviceToHost, wk->s_out));
        return true;
    }
    if (!ensure_slots(wk, false)) return false;
    const size_t col_step = std::min(width, kSlotBytes);
    for (size_t col = 0; col < width; col += col_step) {
        const size_t wd = std::min(col_step, width - col);
        const size_t rows_per = std::max<size_t>(1, kSlotBytes / wd);
        for (size_t r0 = 0; r0 < m.rows; r0 += rows_per) {
            const size_t nr = std::min(rows_per, m.rows - r0);
            if ((int)wk->pending.size() == kSlots && !drain_one(wk)) return false;
            const int slot = wk->next_out;
            wk->next_out = (slot + 1) % kSlots;
            if (nr == 1)
                MMS_CUDA_P(this, cudaMemcpyAsync(wk->pin_out[slot], m.dev + r0 * m.dev_pitch + off + col, wd,
                                                 cudaMemcpyDeviceToHost, wk->s_out));
            else
                MMS_CUDA_P(this, cudaMemcpy2DAsync(wk->pin_out[slot], wd, m.dev + r0 * m.dev_pitch + off + col, m.dev_pitch, wd, nr,
                                                   cudaMemcpyDeviceToHost, wk->s_out));
            MMS_CUDA_P(this, cudaEventRecord(wk->ev_slot_out[slot], wk->s_out));
            wk->pending.push_back({slot, m.host + r0 * m.host_pitch + off + col, m.host_pitch, wd, nr});
        }
    }
    return true;
}

// this worker's inputs of chunk k are on their way: the last worker to say so enqueues the chunk's kernels
void Pipe::arrive(Worker *wk, const PipeJob &job, int k)
{
    PipeChunk &c = chunks[k];
    const int total = team_size(job);
    if (!failed_.load()) {
        if (cudaEventRecord(c.ev_in[wk->id], wk->s_in) != cudaSuccess) set_error(MMS_ERR_CUDA, "cudaEventRecord failed");
    }
    if (arrived_[k].fetch_add(1, std::memory_order_acq_rel) + 1 != total) return;
    bool ok = !failed_.load();
    if (ok) {
        for (int w = 0; w < total && ok; ++w)
            if (cudaStreamWaitEvent(s_compute, c.ev_in[w], 0) != cudaSuccess) { set_error(MMS_ERR_CUDA, "cudaStreamWaitEvent failed"); ok = false; }
    }
    if (ok && job.inputs_ready) ok = job.inputs_ready(k);
    if (ok && job.kind == PIPE_RUN) {
        ok = cudaEventRecord(c.ev_t0, s_compute) == cudaSuccess;
        if (ok && job.run_chunk) ok = job.run_chunk(k);
        if (ok) ok = cudaEventRecord(c.ev_t1, s_compute) == cudaSuccess && cudaEventRecord(c.ev_kernel, s_compute) == cudaSuccess;
        if (!ok) set_error(MMS_ERR_CUDA, "launching the chunk's kernels failed");
    }
    {
        std::lock_guard<std::mutex> lk(mu_);
        enqueued_[k] = 1;
        if (ok && job.kind == PIPE_RUN) c.kernel_recorded = true;
    }
    cv_.notify_all();
}

bool Pipe::do_in(Worker *wk, const PipeJob &job, int k)
{
    const PipeChunk &c = chunks[k];
    int64_t w0 = 0, wn = 0;
    worker_window(c, wk->id, team_size(job), window_unit, w0, wn);
    if (wk->id == 0 && k == 0 && job.before_first_copy && !job.before_first_copy(wk->s_in)) return false;
    for (const Mat &m : job.in)
        if (!stage_in(wk, m, w0, wn)) return false;
    return true;
}

bool Pipe::do_out(Worker *wk, const PipeJob &job, int k)
{
    PipeChunk &c = chunks[k];
    if (job.kind == PIPE_RUN) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return enqueued_[k] || failed_.load(); });
    }
    if (failed_.load()) return false;
    if (c.kernel_recorded) MMS_CUDA_P(this, cudaStreamWaitEvent(wk->s_out, c.ev_kernel, 0));
    int64_t w0 = 0, wn = 0, o0 = 0, on = 0;
    worker_window(c, wk->id, team_size(job), window_unit, w0, wn);
    if (job.out_window) job.out_window(w0, wn, o0, on);
    else { o0 = w0; on = wn; }
    for (const Mat &m : job.out)
        if (!stage_out(wk, m, o0, on)) return false;
    while (!wk->pending.empty())
        if (!drain_one(wk)) return false;
    return true;
}

void Pipe::run_job(Worker *wk, const PipeJob &job)
{
    const int n = (int)chunks.size();
    // workers beyond the job's team (pinned buffers need few copy streams, small jobs few stagers) sit this one out
    if (job.kind != PIPE_CALL && wk->id >= team_size(job)) return;
    switch (job.kind) {
    case PIPE_CALL:
        if (wk->id == 0 && job.call && !job.call() && !failed_.load()) set_error(MMS_ERR_CUDA, "device set-up failed");
        break;
    case PIPE_UPLOAD:
    case PIPE_RUN: {
        bool ok = true;
        for (int k = 0; k <= n; ++k) {
            if (k < n) {
                if (ok && !failed_.load()) ok = do_in(wk, job, k);
                arrive(wk, job, k);              // always, so that nobody waits for a worker that gave up
            }
            if (job.kind == PIPE_RUN && k >= 1 && ok && !failed_.load()) ok = do_out(wk, job, k - 1);
        }
        // the host buffers may be reused as soon as the call returns
        if (cudaStreamSynchronize(wk->s_in) != cudaSuccess) set_error(MMS_ERR_CUDA, "cudaStreamSynchronize failed");
        if (job.kind == PIPE_RUN && cudaStreamSynchronize(wk->s_out) != cudaSuccess) set_error(MMS_ERR_CUDA, "cudaStreamSynchronize failed");
        break;
    }
    case PIPE_DOWNLOAD: {
        bool ok = true;
        for (int k = 0; k < n && ok && !failed_.load(); ++k) ok = do_out(wk, job, k);
        if (cudaStreamSynchronize(wk->s_out) != cudaSuccess) set_error(MMS_ERR_CUDA, "cudaStreamSynchronize failed");
        break;
    }
    default:
        break;
    }
    wk->pending.clear();
}

void Pipe::worker_main(Worker *wk)
{
    bool ok = cudaSetDevice(device) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&wk->s_in, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&wk->s_out, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) set_error(MMS_ERR_CUDA, "a host worker could not create its streams");
    uint64_t seen = 0;
    for (;;) {
        const PipeJob *job = nullptr;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return job_seq_ != seen; });
            seen = job_seq_;
            job = job_;
        }
        if (job->kind == PIPE_EXIT) break;
        if (ok) run_job(wk, *job);
        else set_error(MMS_ERR_CUDA, "a host worker could not create its streams");
        {
            std::lock_guard<std::mutex> lk(mu_);
            pending_ -= 1;
        }
        cv_.notify_all();
    }
    for (int k = 0; k < kSlots; ++k) {
        pinned_slot_release(wk->pin_in[k]);
        pinned_slot_release(wk->pin_out[k]);
        if (wk->ev_slot_in[k]) cudaEventDestroy(wk->ev_slot_in[k]);
        if (wk->ev_slot_out[k]) cudaEventDestroy(wk->ev_slot_out[k]);
    }
    if (wk->s_in) cudaStreamDestroy(wk->s_in);
    if (wk->s_out) cudaStreamDestroy(wk->s_out);
}

}  // namespace mms_api
