// Fiber scheduler behind tests/emul/cuda_emul.h (TEST INFRASTRUCTURE ONLY, see that header).
#include "cuda_emul.h"
#include <chrono>
#include <stdexcept>

namespace emul {
uint3_ g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;
unsigned char* g_smem = nullptr;

namespace {
constexpr size_t kStack = 256 * 1024;
struct Fiber { ucontext_t ctx; std::vector<unsigned char> stack; bool done = false; };
std::vector<Fiber> fibers;
ucontext_t sched_ctx;
int cur = -1;
const std::function<void()>* cur_body = nullptr;
std::vector<double> xch;
std::vector<long long> xch_ll;

void trampoline() {
    (*cur_body)();
    fibers[cur].done = true;
    swapcontext(&fibers[cur].ctx, &sched_ctx);
}
}  // namespace

// A barrier is simply "yield to the scheduler": the scheduler resumes every live fiber once
// per sweep, so when this fiber runs again every other fiber has reached its own barrier
// (all kernels call barriers in block-uniform control flow).
void sync() { swapcontext(&fibers[cur].ctx, &sched_ctx); }

double shfl_xor(double v, int m) {
    xch[cur] = v;
    sync();
    int lane = cur & 31, base = cur & ~31;
    int src = base + (lane ^ m);
    double r = (src < (int)fibers.size()) ? xch[src] : v;
    sync();
    return r;
}
long long shfl_xor_ll(long long v, int m) {
    xch_ll[cur] = v;
    sync();
    int lane = cur & 31, base = cur & ~31;
    int src = base + (lane ^ m);
    long long r = (src < (int)fibers.size()) ? xch_ll[src] : v;
    sync();
    return r;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > 1024) throw std::runtime_error("emul: bad block size");
    std::vector<unsigned char> smem(smem_bytes + 1024);
    g_smem = smem.data();
    g_blockDim = block; g_gridDim = grid;
    if ((int)fibers.size() < nthreads) fibers.resize(nthreads);
    xch.assign(nthreads, 0.0); xch_ll.assign(nthreads, 0);
    cur_body = &body;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        for (int t = 0; t < nthreads; ++t) {
            Fiber& f = fibers[t];
            if (f.stack.size() != kStack) f.stack.resize(kStack);
            f.done = false;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack.data();
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = &sched_ctx;
            makecontext(&f.ctx, trampoline, 0);
        }
        int live = nthreads;
        while (live > 0) {
            live = 0;
            for (int t = 0; t < nthreads; ++t) {
                if (fibers[t].done) continue;
                cur = t;
                g_blockIdx = {bx, by, bz};
                g_threadIdx = {(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                swapcontext(&sched_ctx, &fibers[t].ctx);
                if (!fibers[t].done) ++live;
            }
        }
    }
    cur = -1;
    g_smem = nullptr;
}
}  // namespace emul

// ---- shared-memory backed "device" allocations + IPC handles ------------------------------
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <map>
#include <string>
namespace {
struct ShmBlock { std::string name; size_t bytes; bool owner; };
std::map<void*, ShmBlock> g_blocks;
int g_counter = 0;
}
cudaError_t emul_shm_malloc(void** p, size_t n) {
    if (n == 0) n = 1;
    char name[64];
    snprintf(name, sizeof(name), "/qr_emul_%d_%d", (int)getpid(), g_counter++);
    int fd = shm_open(name, O_CREAT | O_RDWR | O_EXCL, 0600);
    if (fd < 0) { *p = nullptr; return cudaErrorMemoryAllocation; }
    if (ftruncate(fd, (off_t)n) != 0) { close(fd); shm_unlink(name); *p = nullptr; return cudaErrorMemoryAllocation; }
    void* q = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (q == MAP_FAILED) { shm_unlink(name); *p = nullptr; return cudaErrorMemoryAllocation; }
    g_blocks[q] = ShmBlock{name, n, true};
    *p = q;
    return cudaSuccess;
}
cudaError_t emul_shm_free(void* p) {
    if (!p) return cudaSuccess;
    auto it = g_blocks.find(p);
    if (it == g_blocks.end()) return cudaErrorInvalidValue;
    munmap(p, it->second.bytes);
    if (it->second.owner) shm_unlink(it->second.name.c_str());
    g_blocks.erase(it);
    return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
    auto it = g_blocks.find(p);
    if (it == g_blocks.end()) return cudaErrorInvalidValue;
    memset(h->reserved, 0, sizeof(h->reserved));
    strncpy(h->reserved, it->second.name.c_str(), sizeof(h->reserved) - 1);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
    int fd = shm_open(h.reserved, O_RDWR, 0600);
    if (fd < 0) return cudaErrorInvalidValue;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return cudaErrorInvalidValue; }
    void* q = mmap(nullptr, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (q == MAP_FAILED) return cudaErrorInvalidValue;
    g_blocks[q] = ShmBlock{h.reserved, (size_t)st.st_size, false};
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) { return emul_shm_free(p); }

double emul_now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
