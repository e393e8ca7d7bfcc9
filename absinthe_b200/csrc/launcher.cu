// absinthe_b200 launcher: the C ABI of include/absinthe_b200.h.
//
// One shared library, built once (see __graft_entry__.build), drives every generated variant.
// A variant arrives as (descriptor text, sm_100a cubin): the descriptor lists the global fields,
// the fused-group kernels with their launch shapes and parameter-block layout (TMA tensor maps
// and field pointers), and the fields that need a periodic halo update after each group.  The
// launcher loads the cubin through the driver API (entry points are resolved at run time with
// cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda and can be
// dlopen'ed on a machine without a GPU), encodes the tensor maps against the bound buffers and
// replays the group sequence -- optionally as one CUDA graph -- on the caller's stream.
//
// Reference lines each piece replaces are cited in include/absinthe_b200.h.

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/absinthe_b200.h"

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code ? code : -1;
}

#define RT(call)                                                                           \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail((int)e_, "%s failed: %s", #call, cudaGetErrorString(e_));          \
    } while (0)

// ---- driver API, resolved lazily ---------------------------------------------------------
struct Driver {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*OccupancyMaxActiveClusters)(int*, CUfunction, const CUlaunchConfig*) = nullptr;   // optional
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    // stream memory operations (optional: the exchange falls back to kernels without them)
    CUresult (*StreamWaitValue64)(CUstream, CUdeviceptr, cuuint64_t, unsigned) = nullptr;
    CUresult (*StreamWriteValue64)(CUstream, CUdeviceptr, cuuint64_t, unsigned) = nullptr;
    bool ready = false;
};
Driver g_drv;

template <typename F>
int resolve(const char* name, F& slot) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult status;
    cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &status);
    if (e != cudaSuccess || status != cudaDriverEntryPointSuccess || fn == nullptr)
        return fail((int)e, "driver entry point %s unavailable (%s)", name, cudaGetErrorString(e));
    slot = reinterpret_cast<F>(fn);
    return 0;
}

int driver_resolve_all() {
    int rc = 0;
    if ((rc = resolve("cuModuleLoadData", g_drv.ModuleLoadData))) return rc;
    if ((rc = resolve("cuModuleUnload", g_drv.ModuleUnload))) return rc;
    if ((rc = resolve("cuModuleGetFunction", g_drv.ModuleGetFunction))) return rc;
    if ((rc = resolve("cuFuncSetAttribute", g_drv.FuncSetAttribute))) return rc;
    if ((rc = resolve("cuOccupancyMaxActiveBlocksPerMultiprocessor",
                      g_drv.OccupancyMaxActiveBlocksPerMultiprocessor))) return rc;
    if ((rc = resolve("cuLaunchKernel", g_drv.LaunchKernel))) return rc;
    if ((rc = resolve("cuTensorMapEncodeTiled", g_drv.TensorMapEncodeTiled))) return rc;
    if ((rc = resolve("cuGetErrorString", g_drv.GetErrorString))) return rc;
    if (resolve("cuOccupancyMaxActiveClusters", g_drv.OccupancyMaxActiveClusters)) g_drv.OccupancyMaxActiveClusters = nullptr;
    if (resolve("cuStreamWaitValue64", g_drv.StreamWaitValue64) || resolve("cuStreamWriteValue64", g_drv.StreamWriteValue64))
        g_drv.StreamWaitValue64 = nullptr, g_drv.StreamWriteValue64 = nullptr;
    g_error.clear();
    return 0;
}

// several host threads may load their first variant at the same time (one per device): the table is
// resolved exactly once; a failure is remembered and reported to every caller
int driver() {
    static std::once_flag once;
    static int status = 0;
    static std::string message;
    std::call_once(once, [] {
        status = driver_resolve_all();
        if (status) message = g_error;
        g_drv.ready = status == 0;
    });
    if (status) g_error = message;
    return status;
}

const char* drv_error(CUresult r) {
    const char* msg = nullptr;
    if (g_drv.GetErrorString && g_drv.GetErrorString(r, &msg) == CUDA_SUCCESS && msg) return msg;
    return "unknown driver error";
}

#define DRV(call)                                                                          \
    do {                                                                                   \
        CUresult r_ = (call);                                                              \
        if (r_ != CUDA_SUCCESS) return fail((int)r_, "%s failed: %s", #call, drv_error(r_)); \
    } while (0)

// ---- generic device kernels ----------------------------------------------------------------

struct Geometry {
    int X, Y, Z, HX, HY, HZ;
    __host__ __device__ long row() const { return X + 2L * HX; }
    __host__ __device__ long rows() const { return Y + 2L * HY; }
    __host__ __device__ long planes() const { return Z + 2L * HZ; }
    __host__ __device__ long plane() const { return row() * rows(); }
    __host__ __device__ long cells() const { return plane() * planes(); }
};

constexpr int kMaxHaloFields = 16;
constexpr int kHostDepth = 2;        // asynchronous host steps one variant keeps in flight
struct FieldList {
    double* ptr[kMaxHaloFields];
    int count;
};

__device__ __forceinline__ int wrap(int p, int h, int s) {
    // padded index -> the interior padded index it mirrors (needs s >= h)
    return p < h ? p + s : (p >= h + s ? p - s : p);
}

// Closed form of make_periodic (ref template_tiling.cpp:136-159): after the i, j, k passes every
// halo cell holds the interior value at the wrapped coordinates.  One launch updates all outputs
// of a group.  mask: bit0 = i faces, bit1 = j faces, bit2 = k faces.  Valid when each extent is at
// least as large as its halo.
__global__ void periodic_kernel(FieldList fields, Geometry g, int mask) {
    const long row = g.row(), rows = g.rows(), plane = g.plane();
    // three slabs of halo cells, enumerated i-fastest
    const long nk = (mask & 4) ? 2L * g.HZ * plane : 0;                    // k faces: full planes
    const long nj = (mask & 2) ? (long)g.Z * 2 * g.HY * row : 0;          // j faces: interior k
    const long ni = (mask & 1) ? (long)g.Z * g.Y * 2 * g.HX : 0;          // i faces: interior j,k
    const long total = nk + nj + ni;
    for (long n = blockIdx.x * (long)blockDim.x + threadIdx.x; n < total;
         n += (long)gridDim.x * blockDim.x) {
        int i, j, k;
        if (n < nk) {
            long kk = n / plane, r = n % plane;
            k = kk < g.HZ ? (int)kk : (int)(kk + g.Z);
            j = (int)(r / row);
            i = (int)(r % row);
        } else if (n < nk + nj) {
            long m = n - nk;
            long per_k = 2L * g.HY * row;
            k = g.HZ + (int)(m / per_k);
            long r = m % per_k;
            long jj = r / row;
            j = jj < g.HY ? (int)jj : (int)(jj + g.Y);
            i = (int)(r % row);
        } else {
            long m = n - nk - nj;
            long per_k = (long)g.Y * 2 * g.HX;
            k = g.HZ + (int)(m / per_k);
            long r = m % per_k;
            j = g.HY + (int)(r / (2 * g.HX));
            long ii = r % (2 * g.HX);
            i = ii < g.HX ? (int)ii : (int)(ii + g.X);
        }
        int si = (mask & 1) ? wrap(i, g.HX, g.X) : i;
        int sj = (mask & 2) ? wrap(j, g.HY, g.Y) : j;
        int sk = (mask & 4) ? wrap(k, g.HZ, g.Z) : k;
        const long dst = i + j * row + k * plane;
        const long src = si + sj * row + sk * plane;
        for (int f = 0; f < fields.count; ++f) fields.ptr[f][dst] = fields.ptr[f][src];
    }
}

// Literal three-pass version for degenerate domains (an extent smaller than its halo), where the
// sequential copies of the reference read cells written earlier in the same pass.
__global__ void periodic_pass_kernel(double* data, Geometry g, int pass) {
    const long row = g.row(), rows = g.rows(), planes = g.planes(), plane = g.plane();
    const long lines = pass == 0 ? rows * planes : (pass == 1 ? row * planes : row * rows);
    for (long n = blockIdx.x * (long)blockDim.x + threadIdx.x; n < lines;
         n += (long)gridDim.x * blockDim.x) {
        if (pass == 0) {
            double* p = data + n * row;
            for (int i = 0; i < g.HX; ++i) { p[i] = p[i + g.X]; p[i + g.X + g.HX] = p[i + g.HX]; }
        } else if (pass == 1) {
            double* p = data + (n / row) * plane + (n % row);
            for (int j = 0; j < g.HY; ++j) {
                p[j * row] = p[(j + g.Y) * row];
                p[(j + g.Y + g.HY) * row] = p[(j + g.HY) * row];
            }
        } else {
            double* p = data + n;
            for (int k = 0; k < g.HZ; ++k) {
                p[k * plane] = p[(k + g.Z) * plane];
                p[(k + g.Z + g.HZ) * plane] = p[(k + g.HZ) * plane];
            }
        }
    }
}

__global__ void flush_kernel(double* buf, long n, double v) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        buf[i] = v;
}

// rows [j0, j0+width) of every plane <-> dense [planes][width][row]
__global__ void pack_kernel(const double* __restrict__ field, double* __restrict__ buffer, Geometry g,
                            int j0, int width, int unpack) {
    const long row = g.row(), plane = g.plane();
    const long total = g.planes() * width * row;
    for (long n = blockIdx.x * (long)blockDim.x + threadIdx.x; n < total;
         n += (long)gridDim.x * blockDim.x) {
        long k = n / (width * row), r = n % (width * row);
        long at = k * plane + (j0 + r / row) * row + r % row;
        if (unpack) const_cast<double*>(field)[at] = buffer[n];
        else buffer[n] = field[at];
    }
}

// both j-faces of every field of a list in one launch: buffers are [field][plane][HY][row];
// pack reads the first / last HY interior rows, unpack writes the low / high halo rows
__global__ void faces_kernel(FieldList fields, double* low, double* high, Geometry g, int unpack) {
    const long row = g.row(), plane = g.plane();
    const long per_field = g.planes() * g.HY * row;
    const long total = 2 * fields.count * per_field;
    for (long n = blockIdx.x * (long)blockDim.x + threadIdx.x; n < total;
         n += (long)gridDim.x * blockDim.x) {
        const int side = (int)(n / (fields.count * per_field));
        const long m = n % (fields.count * per_field);
        const int f = (int)(m / per_field);
        const long r = m % per_field;
        const long k = r / (g.HY * row), q = r % (g.HY * row);
        // pack: low side = rows [HY, 2HY), high side = rows [Y, Y+HY)
        // unpack: low buffer -> rows [0, HY), high buffer -> rows [Y+HY, Y+2HY)
        const long j0 = unpack ? (side ? g.Y + g.HY : 0) : (side ? g.Y : g.HY);
        const long at = k * plane + (j0 + q / row) * row + q % row;
        double* buffer = side ? high : low;
        if (unpack) fields.ptr[f][at] = buffer[m];
        else buffer[m] = fields.ptr[f][at];
    }
}

constexpr int kMaxDevices = 64;

// multiprocessors of the current device (queried once per device)
int sm_count() {
    static int cached[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 132;
    if (!cached[dev]) {
        int n = 0;
        cached[dev] = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0 ? n : 132;
    }
    return cached[dev];
}

// ---- P2P slab exchange -------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long load_flag(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void store_flag(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// The exchange kernels are built to squeeze in beside a compute kernel: 128 threads, one row per block pass,
// 16-byte accesses, no 64-bit divisions (a fused group kernel can leave less than 4000 registers per SM free).
constexpr int kExchangeThreads = 128;

// rows [j0, j0+HY) of plane k of field f <-> buffer[((f * planes + k) * HY + r) * row ...]
__device__ __forceinline__ void move_rows(FieldList fields, const Geometry& g, double* buffer, int j0, bool to_buffer) {
    const int row = (int)g.row(), planes = (int)g.planes();
    const long plane = g.plane();
    const int lines = fields.count * planes * g.HY;
    for (int line = blockIdx.x; line < lines; line += gridDim.x) {
        const int r = line % g.HY, k = (line / g.HY) % planes, f = line / (g.HY * planes);
        double* field = fields.ptr[f] + k * plane + (long)(j0 + r) * row;
        double* dense = buffer + (long)line * row;
        if ((row & 1) == 0) {                                   // even rows: both sides are 16-byte aligned
            double2* a = reinterpret_cast<double2*>(to_buffer ? dense : field);
            const double2* b = reinterpret_cast<const double2*>(to_buffer ? field : dense);
            for (int i = threadIdx.x; i < row / 2; i += blockDim.x) a[i] = b[i];
        } else {
            for (int i = threadIdx.x; i < row; i += blockDim.x) (to_buffer ? dense : field)[i] = (to_buffer ? field : dense)[i];
        }
    }
}

// PUT: the first HY interior rows of every field go to the previous rank's "from next" buffer, the last HY
// interior rows to the next rank's "from previous" buffer ([field][plane][HY][row], through peer-mapped
// pointers); the last block to finish raises both arrival flags to `epoch`.
__global__ void __launch_bounds__(kExchangeThreads)
put_kernel(FieldList fields, Geometry g, double* to_prev, double* to_next, unsigned long long* flag_at_prev,
           unsigned long long* flag_at_next, unsigned* done, unsigned long long epoch) {
    move_rows(fields, g, to_prev, g.HY, true);
    move_rows(fields, g, to_next, g.Y, true);
    __threadfence_system();                                 // this thread's remote stores are out
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(done, 1u) == gridDim.x - 1) {         // every block's stores are out
            *done = 0;
            __threadfence_system();
            store_flag(flag_at_prev, epoch);
            store_flag(flag_at_next, epoch);
        }
    }
}

// raise flags from a kernel (copy-engine PUT on a driver without stream memory operations)
__global__ void signal_kernel(unsigned long long* flag_at_prev, unsigned long long* flag_at_next, unsigned long long epoch) {
    __threadfence_system();
    store_flag(flag_at_prev, epoch);
    store_flag(flag_at_next, epoch);
}

// spin (bounded) until both neighbours have raised this rank's flags to `epoch`
__global__ void await_kernel(const unsigned long long* flags, unsigned long long epoch) {
    unsigned long long spins = 0;
    while (load_flag(flags) < epoch || load_flag(flags + 1) < epoch) {
        __nanosleep(200);
        if (++spins > (1ull << 24)) __trap();               // a neighbour that never arrives ends in an error, not a hang
    }
}

// the rows the neighbours stored -> low / high j-halo
__global__ void __launch_bounds__(kExchangeThreads)
unpack_kernel(FieldList fields, Geometry g, double* from_prev, double* from_next) {
    move_rows(fields, g, from_prev, 0, false);
    move_rows(fields, g, from_next, g.Y + g.HY, false);
}

int launch_grid(long items, int threads) {
    long blocks = (items + threads - 1) / threads;
    return (int)std::max(1L, std::min(blocks, (long)sm_count() * 16));
}

int periodic_update(const FieldList& fields, const Geometry& g, int mask, cudaStream_t stream) {
    if (fields.count == 0) return 0;
    const bool regular = g.X >= g.HX && g.Y >= g.HY && g.Z >= g.HZ;
    if (regular) {
        long total = 2L * g.HZ * g.plane() + (long)g.Z * 2 * g.HY * g.row() + (long)g.Z * g.Y * 2 * g.HX;
        periodic_kernel<<<launch_grid(total, 256), 256, 0, stream>>>(fields, g, mask);
    } else {
        for (int f = 0; f < fields.count; ++f)
            for (int pass = 0; pass < 3; ++pass) {
                if (!(mask & (1 << pass))) continue;
                long lines = pass == 0 ? g.rows() * g.planes()
                                       : (pass == 1 ? g.row() * g.planes() : g.row() * g.rows());
                periodic_pass_kernel<<<launch_grid(lines, 128), 128, 0, stream>>>(fields.ptr[f], g, pass);
            }
    }
    RT(cudaGetLastError());
    return 0;
}

// ---- variant -------------------------------------------------------------------------------

enum FieldKind { kInput = 0, kOutput = 1, kTemp = 2 };

struct Field {
    std::string name;
    FieldKind kind;
    double* ptr = nullptr;
    bool owned = false;
};

struct Arg {
    enum Type { kMap, kPtr, kTileCtl, kScratch, kRemote } type;
    size_t offset;
    int field = -1;
    unsigned box[3] = {0, 0, 0};
};

struct Group {
    int id = 0;
    std::string symbol;
    int threads = 0;
    size_t smem = 0;
    int tiles = 0;
    size_t params_bytes = 0;
    size_t scratch_bytes = 0;      // per resident CTA
    int training_step = 20;
    int fused_halo = 0;
    int cluster = 1;               // CTAs per thread-block cluster (compiled into the kernel as __cluster_dims__)
    int max_clusters = 0;          // co-resident clusters of the device (0: not a cluster kernel)
    std::vector<Arg> args;
    std::vector<int> halo_fields;
    CUfunction fn = nullptr;
    int blocks_per_sm = 1;
    std::vector<unsigned char> params;
    double* scratch = nullptr;
};

// interior comparison of the reference's verifier (ref template_tiling.cpp:383-392): a point is an error
// when |expected - actual| > max(|expected|, |actual|) * tolerance, or when it is not finite although expected is
__global__ void compare_kernel(const double* __restrict__ expected, const double* __restrict__ actual,
                               Geometry g, double tolerance, unsigned long long* counters) {
    const long cells = (long)g.X * g.Y * g.Z;
    unsigned long long errors = 0, matches = 0;
    for (long n = blockIdx.x * (long)blockDim.x + threadIdx.x; n < cells; n += (long)gridDim.x * blockDim.x) {
        const long i = n % g.X, j = (n / g.X) % g.Y, k = n / ((long)g.X * g.Y);
        const long at = (i + g.HX) + (j + g.HY) * g.row() + (k + g.HZ) * g.plane();
        const double e = expected[at], a = actual[at];
        // the reference's test never fires on NaN (every comparison with NaN is false, ref :386): a variant
        // that returns NaN or inf where the reference is finite must not pass as "0 errors"
        if (fabs(e - a) > fmax(fabs(e), fabs(a)) * tolerance || (isfinite(e) && !isfinite(a))) ++errors;
        else ++matches;
    }
    for (int d = 16; d > 0; d >>= 1) {
        errors += __shfl_down_sync(0xffffffffu, errors, d);
        matches += __shfl_down_sync(0xffffffffu, matches, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (errors) atomicAdd(&counters[0], errors);
        if (matches) atomicAdd(&counters[1], matches);
    }
}

}  // namespace

struct absinthe_variant {
    std::string name;
    Geometry geo{};
    int device = 0;
    int sms = 148;
    std::vector<Field> fields;
    std::vector<int> inputs, outputs;
    std::vector<Group> groups;
    CUmodule module = nullptr;
    bool bound = false;
    cudaGraphExec_t graph = nullptr;
    cudaStream_t graph_stream = nullptr;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_done = nullptr;
    cudaEvent_t ev_out[kHostDepth] = {};                  // outputs of async host step n are in ev_out[n % depth]
    unsigned long long host_issued = 0, host_waited = 0;  // async host steps enqueued / waited for
    bool host_interior = false;                           // upload interiors only, rebuild the halos on the device
    std::vector<cudaEvent_t> events;
    // P2P slab exchange (absinthe_comm_*): this rank's window and the neighbours' windows
    struct CommGroup {
        int group = 0;
        size_t face_doubles = 0;                 // one direction of one PUT: fields x planes x HY x row
        size_t data = 0;                         // offset (doubles) of [parity][from_prev | from_next] buffers
        size_t flags = 0;                        // offset (doubles) of {from_prev, from_next} arrival epochs
        unsigned long long put = 0, waited = 0;  // PUTs enqueued / WAITs enqueued
        int engine = 0;                          // engine of the outstanding PUT
        cudaEvent_t unpacked = nullptr;          // recorded after the WAIT kernel of the latest epoch
    };
    int comm_rank = 0, comm_ranks = 0;
    double* window = nullptr;                    // own window (cudaMalloc, IPC exported)
    size_t window_doubles = 0;
    double* window_prev = nullptr;               // neighbours' windows (IPC mapped, or own window when n = 1)
    double* window_next = nullptr;
    bool prev_mapped = false, next_mapped = false;
    unsigned* put_done = nullptr;                // block counter of the PUT kernels (one per group)
    cudaStream_t unpack_stream = nullptr;        // WAITs run here; the caller's stream waits for their event
    std::vector<CommGroup> comm_groups;
};

namespace {

std::mutex g_device_mutex;                       // guards the per-device scratch below
double* g_flush[kMaxDevices] = {};               // one flush buffer per device
unsigned long long* g_counters[kMaxDevices] = {};   // error / match counters of absinthe_compare_fields
const long kFlushDoubles = 48L * 1024 * 1024;    // 384 MB > 2 x L2

// scratch of the current device, allocated on first use
template <typename T>
int device_scratch(T* (&table)[kMaxDevices], size_t bytes, T** out) {
    int dev = 0;
    RT(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(-3, "device %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_device_mutex);
    if (!table[dev]) RT(cudaMalloc(&table[dev], bytes));
    *out = table[dev];
    return 0;
}

// every entry point that launches runs on the variant's device, whatever device the caller left current
struct OnDevice {
    int previous = -1;
    bool ok = true;
    explicit OnDevice(int device) {
        if (cudaGetDevice(&previous) != cudaSuccess) previous = -1;
        if (previous != device) ok = cudaSetDevice(device) == cudaSuccess;
    }
    ~OnDevice() {
        int now = -1;
        if (previous >= 0 && cudaGetDevice(&now) == cudaSuccess && now != previous) cudaSetDevice(previous);
    }
};

int find_field(const absinthe_variant* v, const std::string& name) {
    for (size_t n = 0; n < v->fields.size(); ++n)
        if (v->fields[n].name == name) return (int)n;
    return -1;
}

int parse_descriptor(absinthe_variant* v, const char* text) {
    std::istringstream in(text);
    std::string line;
    Group* group = nullptr;
    bool header = false;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string key;
        if (!(ls >> key) || key[0] == '#') continue;
        if (key == "absinthe-variant") {
            int version = 0;
            ls >> version;
            if (version != ABSINTHE_ABI_VERSION) return fail(-2, "descriptor version %d unsupported", version);
            header = true;
        } else if (key == "name") {
            ls >> v->name;
        } else if (key == "domain") {
            ls >> v->geo.X >> v->geo.Y >> v->geo.Z >> v->geo.HX >> v->geo.HY >> v->geo.HZ;
        } else if (key == "field") {
            Field f;
            std::string kind;
            ls >> f.name >> kind;
            f.kind = kind == "in" ? kInput : (kind == "out" ? kOutput : kTemp);
            if (f.kind == kInput) v->inputs.push_back((int)v->fields.size());
            if (f.kind == kOutput) v->outputs.push_back((int)v->fields.size());
            v->fields.push_back(f);
        } else if (key == "group") {
            v->groups.emplace_back();
            group = &v->groups.back();
            std::string tok;
            ls >> group->id;
            while (ls >> tok) {
                if (tok == "kernel") ls >> group->symbol;
                else if (tok == "threads") ls >> group->threads;
                else if (tok == "smem") ls >> group->smem;
                else if (tok == "tiles") ls >> group->tiles;
                else if (tok == "params") ls >> group->params_bytes;
                else if (tok == "scratch") ls >> group->scratch_bytes;
                else if (tok == "training_step") ls >> group->training_step;
                else if (tok == "fused_halo") ls >> group->fused_halo;
                else if (tok == "cluster") ls >> group->cluster;
                else return fail(-2, "descriptor: unknown group attribute '%s'", tok.c_str());
            }
        } else if (key == "arg") {
            if (!group) return fail(-2, "descriptor: arg outside group");
            Arg a{};
            std::string type, name;
            ls >> type >> a.offset;
            if (type == "tmap") {
                a.type = Arg::kMap;
                ls >> name >> a.box[0] >> a.box[1] >> a.box[2];
            } else if (type == "ptr") {
                a.type = Arg::kPtr;
                ls >> name;
            } else if (type == "tilectl") {
                a.type = Arg::kTileCtl;
            } else if (type == "scratch") {
                a.type = Arg::kScratch;
            } else if (type == "remote") {
                a.type = Arg::kRemote;
            } else {
                return fail(-2, "descriptor: unknown arg type '%s'", type.c_str());
            }
            if (!name.empty()) {
                a.field = find_field(v, name);
                if (a.field < 0) return fail(-2, "descriptor: unknown field '%s'", name.c_str());
            }
            group->args.push_back(a);
        } else if (key == "halo") {
            if (!group) return fail(-2, "descriptor: halo outside group");
            std::string name;
            ls >> name;
            int f = find_field(v, name);
            if (f < 0) return fail(-2, "descriptor: unknown halo field '%s'", name.c_str());
            group->halo_fields.push_back(f);
        } else if (key == "endgroup") {
            group = nullptr;
        } else if (key == "end") {
            break;
        }
    }
    if (!header) return fail(-2, "descriptor: missing 'absinthe-variant' header");
    if (v->geo.X <= 0 || v->geo.Y <= 0 || v->geo.Z <= 0) return fail(-2, "descriptor: bad domain");
    for (auto& g : v->groups)
        if ((int)g.halo_fields.size() > kMaxHaloFields)
            return fail(-2, "group %d has more than %d halo fields", g.id, kMaxHaloFields);
    return 0;
}

int encode_map(const absinthe_variant* v, const Arg& a, double* base, CUtensorMap* map) {
    const Geometry& g = v->geo;
    cuuint64_t dims[3] = {(cuuint64_t)g.row(), (cuuint64_t)g.rows(), (cuuint64_t)g.planes()};
    cuuint64_t strides[2] = {(cuuint64_t)g.row() * 8, (cuuint64_t)g.plane() * 8};
    cuuint32_t box[3] = {a.box[0], a.box[1], a.box[2]};
    cuuint32_t elem[3] = {1, 1, 1};
    // L2 promotion: how much the L2 fetches around a requested sector.  256 B suits boxes whose rows are
    // read whole by neighbouring work items soon after; ABSINTHE_TMA_L2_PROMOTION=0..3 (none, 64, 128,
    // 256 B) overrides it for A/B runs.
    CUtensorMapL2promotion promotion = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (const char* env = std::getenv("ABSINTHE_TMA_L2_PROMOTION")) {
        const int level = std::atoi(env);
        promotion = level <= 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                               : (level == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                             : (level == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                           : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
    }
    DRV(g_drv.TensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box,
                                   elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   promotion, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    return 0;
}

int grid_for(const absinthe_variant* v, const Group& g, int tiles) {
    int resident = std::max(1, g.blocks_per_sm) * v->sms;
    if (g.cluster > 1) {
        // a cluster kernel walks super-tiles of `cluster` tiles: the grid is a whole number of clusters, all of
        // them resident at once (a GPC whose SM count is no multiple of the cluster size leaves SMs idle)
        int clusters = g.max_clusters > 0 ? g.max_clusters : resident / g.cluster;
        clusters = std::min(clusters, (tiles + g.cluster - 1) / g.cluster);
        return std::max(1, clusters) * g.cluster;
    }
    return std::max(1, std::min(tiles, resident));
}

int build_params(absinthe_variant* v, Group& g) {
    g.params.assign(g.params_bytes, 0);
    for (const Arg& a : g.args) {
        if (a.type == Arg::kMap) {
            CUtensorMap map;
            int rc = encode_map(v, a, v->fields[a.field].ptr, &map);
            if (rc) return rc;
            std::memcpy(g.params.data() + a.offset, &map, sizeof(map));
        } else if (a.type == Arg::kPtr) {
            double* p = v->fields[a.field].ptr;
            std::memcpy(g.params.data() + a.offset, &p, sizeof(p));
        } else if (a.type == Arg::kScratch) {
            std::memcpy(g.params.data() + a.offset, &g.scratch, sizeof(g.scratch));
        }
    }
    return 0;
}

// neighbours' exchange windows for the j-wrapped copies of a fused-halo kernel (null: the rank's own j-halo)
bool set_remote(Group& g, double* to_prev, double* to_next) {
    for (const Arg& a : g.args)
        if (a.type == Arg::kRemote) {
            double* both[2] = {to_prev, to_next};
            std::memcpy(g.params.data() + a.offset, both, sizeof(both));
            return true;
        }
    return false;
}

void set_tilectl(Group& g, int step, int count) {
    for (const Arg& a : g.args)
        if (a.type == Arg::kTileCtl) {
            int ctl[2] = {step, count};
            std::memcpy(g.params.data() + a.offset, ctl, sizeof(ctl));
        }
}

int launch_group(absinthe_variant* v, Group& g, bool training, cudaStream_t stream) {
    if (training && g.cluster > 1)
        return fail(-4, "kernel %s walks super-tiles of a thread-block cluster: no training loop over single tiles", g.symbol.c_str());
    int step = training ? g.training_step : 1;
    int count = (g.tiles + step - 1) / step;
    set_tilectl(g, step, count);
    void* params[1] = {g.params.data()};
    int grid = grid_for(v, g, count);
    DRV(g_drv.LaunchKernel(g.fn, grid, 1, 1, g.threads, 1, 1, (unsigned)g.smem, (CUstream)stream,
                           params, nullptr));
    return 0;
}

int halo_group(absinthe_variant* v, Group& g, cudaStream_t stream) {
    if (g.fused_halo || g.halo_fields.empty()) return 0;
    FieldList list{};
    list.count = (int)g.halo_fields.size();
    for (int n = 0; n < list.count; ++n) list.ptr[n] = v->fields[g.halo_fields[n]].ptr;
    return periodic_update(list, v->geo, 7, stream);
}

int require_bound(const absinthe_variant* v) {
    if (!v) return fail(-3, "null variant");
    if (!v->bound) return fail(-3, "variant '%s' has no buffers bound", v->name.c_str());
    return 0;
}

}  // namespace

// ---- C ABI ---------------------------------------------------------------------------------

extern "C" {

int absinthe_abi_version(void) { return ABSINTHE_ABI_VERSION; }

const char* absinthe_last_error(void) { return g_error.c_str(); }

int absinthe_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        fail((int)e, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return -(int)e;
    }
    return n;
}

int absinthe_variant_load(const char* descriptor, const void* cubin, size_t cubin_bytes, int device,
                          absinthe_variant** out) {
    if (!descriptor || !cubin || !cubin_bytes || !out) return fail(-1, "absinthe_variant_load: null argument");
    *out = nullptr;
    RT(cudaSetDevice(device));
    RT(cudaFree(nullptr));
    int rc = driver();
    if (rc) return rc;
    absinthe_variant* v = new absinthe_variant();
    v->device = device;
    rc = parse_descriptor(v, descriptor);
    if (rc) { delete v; return rc; }
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete v; return fail((int)e, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    v->sms = prop.multiProcessorCount;
    // failure paths: format the message while the variant (and the strings it owns) is alive, then free it
    auto give_up = [&](int code) { absinthe_variant_free(v); return code; };
    CUresult r = g_drv.ModuleLoadData(&v->module, cubin);
    if (r != CUDA_SUCCESS) return give_up(fail((int)r, "cuModuleLoadData failed: %s (is the cubin built for sm_%d%d?)", drv_error(r), prop.major, prop.minor));
    for (Group& g : v->groups) {
        r = g_drv.ModuleGetFunction(&g.fn, v->module, g.symbol.c_str());
        if (r != CUDA_SUCCESS) return give_up(fail((int)r, "kernel %s not found: %s", g.symbol.c_str(), drv_error(r)));
        if (g.smem > 48 * 1024) {
            r = g_drv.FuncSetAttribute(g.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)g.smem);
            if (r != CUDA_SUCCESS) return give_up(fail((int)r, "kernel %s: %zu bytes of shared memory refused: %s", g.symbol.c_str(), g.smem, drv_error(r)));
        }
        int blocks = 0;
        r = g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&blocks, g.fn, g.threads, g.smem);
        if (r != CUDA_SUCCESS || blocks < 1) return give_up(fail(r != CUDA_SUCCESS ? (int)r : -4, "kernel %s cannot be resident (threads %d, smem %zu)", g.symbol.c_str(), g.threads, g.smem));
        g.blocks_per_sm = blocks;
        if (g.cluster > 1) {
            if (g.cluster > 8) return give_up(fail(-4, "kernel %s: cluster of %d CTAs exceeds the portable limit of 8", g.symbol.c_str(), g.cluster));
            g.max_clusters = std::max(1, blocks * v->sms / g.cluster);
            if (g_drv.OccupancyMaxActiveClusters) {
                CUlaunchAttribute attr{};
                attr.id = CU_LAUNCH_ATTRIBUTE_CLUSTER_DIMENSION;
                attr.value.clusterDim.x = (unsigned)g.cluster;
                attr.value.clusterDim.y = attr.value.clusterDim.z = 1;
                CUlaunchConfig cfg{};
                cfg.gridDimX = (unsigned)(g.max_clusters * g.cluster);
                cfg.gridDimY = cfg.gridDimZ = 1;
                cfg.blockDimX = (unsigned)g.threads;
                cfg.blockDimY = cfg.blockDimZ = 1;
                cfg.sharedMemBytes = (unsigned)g.smem;
                cfg.attrs = &attr;
                cfg.numAttrs = 1;
                int clusters = 0;
                r = g_drv.OccupancyMaxActiveClusters(&clusters, g.fn, &cfg);
                if (r != CUDA_SUCCESS || clusters < 1) return give_up(fail(r != CUDA_SUCCESS ? (int)r : -4, "kernel %s: no cluster of %d CTAs can be resident: %s", g.symbol.c_str(), g.cluster, r != CUDA_SUCCESS ? drv_error(r) : "0 clusters"));
                g.max_clusters = std::min(g.max_clusters, clusters);
            }
        }
        if (g.scratch_bytes) {
            size_t total = g.scratch_bytes * (size_t)grid_for(v, g, g.tiles);
            e = cudaMalloc(&g.scratch, total);
            if (e != cudaSuccess) { g.scratch = nullptr; return give_up(fail((int)e, "scratch allocation of %zu bytes failed", total)); }
            e = cudaMemset(g.scratch, 0, total);
            if (e != cudaSuccess) return give_up(fail((int)e, "scratch of kernel %s: cudaMemset failed: %s", g.symbol.c_str(), cudaGetErrorString(e)));
        }
    }
    // inter-group temporaries are owned by the library (zero-initialised like sarray_3d(0.0))
    for (Field& f : v->fields)
        if (f.kind == kTemp) {
            size_t bytes = (size_t)v->geo.cells() * 8;
            e = cudaMalloc(&f.ptr, bytes);
            if (e != cudaSuccess) { f.ptr = nullptr; return give_up(fail((int)e, "temporary '%s': %zu bytes refused", f.name.c_str(), bytes)); }
            f.owned = true;
            e = cudaMemset(f.ptr, 0, bytes);
            if (e != cudaSuccess) return give_up(fail((int)e, "temporary '%s': cudaMemset failed: %s", f.name.c_str(), cudaGetErrorString(e)));
        }
    *out = v;
    return 0;
}

void absinthe_variant_free(absinthe_variant* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    if (v->graph) cudaGraphExecDestroy(v->graph);
    if (v->graph_stream) cudaStreamDestroy(v->graph_stream);
    if (v->h2d_stream) cudaStreamDestroy(v->h2d_stream);
    if (v->d2h_stream) cudaStreamDestroy(v->d2h_stream);
    for (cudaEvent_t e : {v->ev_in, v->ev_done})
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : v->ev_out)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : v->events) cudaEventDestroy(e);
    absinthe_comm_destroy(v);
    for (Field& f : v->fields)
        if (f.owned && f.ptr) cudaFree(f.ptr);
    for (Group& g : v->groups)
        if (g.scratch) cudaFree(g.scratch);
    if (v->module && g_drv.ModuleUnload) g_drv.ModuleUnload(v->module);
    delete v;
}

int absinthe_variant_info(const absinthe_variant* v, int dims[6], int* n_inputs, int* n_outputs,
                          int* n_groups) {
    if (!v) return fail(-3, "null variant");
    if (dims) {
        dims[0] = v->geo.X; dims[1] = v->geo.Y; dims[2] = v->geo.Z;
        dims[3] = v->geo.HX; dims[4] = v->geo.HY; dims[5] = v->geo.HZ;
    }
    if (n_inputs) *n_inputs = (int)v->inputs.size();
    if (n_outputs) *n_outputs = (int)v->outputs.size();
    if (n_groups) *n_groups = (int)v->groups.size();
    return 0;
}

const char* absinthe_variant_field(const absinthe_variant* v, int is_output, int idx) {
    if (!v) return nullptr;
    const std::vector<int>& list = is_output ? v->outputs : v->inputs;
    if (idx < 0 || idx >= (int)list.size()) return nullptr;
    return v->fields[list[idx]].name.c_str();
}

size_t absinthe_variant_field_bytes(const absinthe_variant* v) {
    return v ? (size_t)v->geo.cells() * 8 : 0;
}

int absinthe_variant_launches(const absinthe_variant* v) {
    if (!v) return 0;
    int n = 0;
    for (const Group& g : v->groups) {
        n += 1;
        if (!g.fused_halo && !g.halo_fields.empty()) {
            const Geometry& q = v->geo;
            bool regular = q.X >= q.HX && q.Y >= q.HY && q.Z >= q.HZ;
            n += regular ? 1 : 3 * (int)g.halo_fields.size();
        }
    }
    return n;
}

int absinthe_variant_fused_halo(const absinthe_variant* v) {
    if (!v) return 0;
    for (const Group& g : v->groups)
        if (!g.fused_halo && !g.halo_fields.empty()) return 0;
    return 1;
}

int absinthe_variant_bind(absinthe_variant* v, const double* const* d_inputs, double* const* d_outputs) {
    if (!v) return fail(-3, "null variant");
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d of variant '%s'", v->device, v->name.c_str());
    for (size_t n = 0; n < v->inputs.size(); ++n) {
        if (!d_inputs || !d_inputs[n]) return fail(-3, "input %zu is null", n);
        v->fields[v->inputs[n]].ptr = const_cast<double*>(d_inputs[n]);
    }
    for (size_t n = 0; n < v->outputs.size(); ++n) {
        if (!d_outputs || !d_outputs[n]) return fail(-3, "output %zu is null", n);
        v->fields[v->outputs[n]].ptr = d_outputs[n];
    }
    for (Group& g : v->groups) {
        int rc = build_params(v, g);
        if (rc) return rc;
    }
    if (v->graph) { cudaGraphExecDestroy(v->graph); v->graph = nullptr; }
    v->bound = true;
    return 0;
}

int absinthe_variant_apply(absinthe_variant* v, void* stream) {
    int rc = require_bound(v);
    if (rc) return rc;
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d of variant '%s'", v->device, v->name.c_str());
    cudaStream_t s = (cudaStream_t)stream;
    for (Group& g : v->groups) {
        if ((rc = launch_group(v, g, false, s))) return rc;
        if ((rc = halo_group(v, g, s))) return rc;
    }
    return 0;
}

int absinthe_variant_apply_graph(absinthe_variant* v, void* stream) {
    int rc = require_bound(v);
    if (rc) return rc;
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d of variant '%s'", v->device, v->name.c_str());
    cudaStream_t s = (cudaStream_t)stream;
    if (!v->graph) {
        // capture on a private stream (the legacy default stream cannot be captured); the
        // instantiated graph is then launched into whatever stream the caller passes
        if (!v->graph_stream) RT(cudaStreamCreateWithFlags(&v->graph_stream, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        RT(cudaStreamBeginCapture(v->graph_stream, cudaStreamCaptureModeThreadLocal));
        rc = absinthe_variant_apply(v, v->graph_stream);
        cudaError_t e = cudaStreamEndCapture(v->graph_stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return fail((int)e, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&v->graph, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail((int)e, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    RT(cudaGraphLaunch(v->graph, s));
    return 0;
}

int absinthe_variant_apply_group(absinthe_variant* v, int group, int with_halo, void* stream) {
    int rc = require_bound(v);
    if (rc) return rc;
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d of variant '%s'", v->device, v->name.c_str());
    for (Group& g : v->groups)
        if (g.id == group) {
            if ((rc = launch_group(v, g, false, (cudaStream_t)stream))) return rc;
            return with_halo ? halo_group(v, g, (cudaStream_t)stream) : 0;
        }
    return fail(-3, "variant '%s' has no group %d", v->name.c_str(), group);
}

int absinthe_variant_halo_group(absinthe_variant* v, int group, void* stream) {
    int rc = require_bound(v);
    if (rc) return rc;
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d of variant '%s'", v->device, v->name.c_str());
    for (Group& g : v->groups)
        if (g.id == group) return halo_group(v, g, (cudaStream_t)stream);
    return fail(-3, "variant '%s' has no group %d", v->name.c_str(), group);
}

static int group_faces(absinthe_variant* v, int group, double* low, double* high, int local_faces, int unpack,
                       void* stream) {
    int rc = require_bound(v);
    if (rc) return rc;
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d of variant '%s'", v->device, v->name.c_str());
    for (Group& g : v->groups)
        if (g.id == group) {
            if (g.halo_fields.empty()) return 0;
            FieldList list{};
            list.count = (int)g.halo_fields.size();
            for (int n = 0; n < list.count; ++n) list.ptr[n] = v->fields[g.halo_fields[n]].ptr;
            const Geometry& q = v->geo;
            if (local_faces && !unpack && !g.fused_halo) {
                if (!(q.X >= q.HX && q.Z >= q.HZ)) return fail(-3, "extent smaller than halo");
                long cells = 2L * q.HZ * q.plane() + (long)q.Z * q.Y * 2 * q.HX;
                periodic_kernel<<<launch_grid(cells, 256), 256, 0, (cudaStream_t)stream>>>(list, q, 5);
            }
            long total = 2L * list.count * q.planes() * q.HY * q.row();
            faces_kernel<<<launch_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(list, low, high, q, unpack);
            RT(cudaGetLastError());
            return 0;
        }
    return fail(-3, "variant '%s' has no group %d", v->name.c_str(), group);
}

int absinthe_variant_pack_group(absinthe_variant* v, int group, double* d_low, double* d_high, int local_faces,
                                void* stream) {
    return group_faces(v, group, d_low, d_high, local_faces, 0, stream);
}

int absinthe_variant_unpack_group(absinthe_variant* v, int group, const double* d_low, const double* d_high,
                                  void* stream) {
    return group_faces(v, group, const_cast<double*>(d_low), const_cast<double*>(d_high), 0, 1, stream);
}

int absinthe_variant_group_outputs(const absinthe_variant* v, int group) {
    if (!v) return 0;
    for (const Group& g : v->groups)
        if (g.id == group) return (int)g.halo_fields.size();
    return 0;
}

const char* absinthe_variant_group_output(const absinthe_variant* v, int group, int idx) {
    if (!v) return nullptr;
    for (const Group& g : v->groups)
        if (g.id == group && idx >= 0 && idx < (int)g.halo_fields.size())
            return v->fields[g.halo_fields[idx]].name.c_str();
    return nullptr;
}

double* absinthe_variant_field_ptr(const absinthe_variant* v, const char* name) {
    if (!v || !name) return nullptr;
    int f = find_field(v, name);
    return f < 0 ? nullptr : v->fields[f].ptr;
}

int absinthe_flush_l2(void* stream) {
    double* buffer = nullptr;
    int rc = device_scratch(g_flush, kFlushDoubles * 8, &buffer);
    if (rc) return rc;
    flush_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(buffer, kFlushDoubles, 0.0);
    RT(cudaGetLastError());
    return 0;
}

int absinthe_variant_run(absinthe_variant* v, void* stream, int runs, int flush, int training,
                         float* total_ms, float* halo_ms) {
    int rc = require_bound(v);
    if (rc) return rc;
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d of variant '%s'", v->device, v->name.c_str());
    if (runs <= 0 || !total_ms || !halo_ms) return fail(-3, "absinthe_variant_run: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t need = 1 + 2 * v->groups.size();
    while (v->events.size() < need) {
        cudaEvent_t e;
        RT(cudaEventCreate(&e));
        v->events.push_back(e);
    }
    for (int run = 0; run < 2 * runs; ++run) {
        if (flush && (rc = absinthe_flush_l2(s))) return rc;
        size_t ev = 0;
        RT(cudaEventRecord(v->events[ev++], s));
        for (Group& g : v->groups) {
            if ((rc = launch_group(v, g, training != 0, s))) return rc;
            RT(cudaEventRecord(v->events[ev++], s));
            if (!training && (rc = halo_group(v, g, s))) return rc;
            RT(cudaEventRecord(v->events[ev++], s));
        }
        RT(cudaStreamSynchronize(s));
        if (run % 2 == 1) {
            float total = 0.f, halo = 0.f;
            RT(cudaEventElapsedTime(&total, v->events[0], v->events[ev - 1]));
            for (size_t g = 0; g < v->groups.size(); ++g) {
                float t = 0.f;
                RT(cudaEventElapsedTime(&t, v->events[1 + 2 * g], v->events[2 + 2 * g]));
                halo += t;
            }
            total_ms[run / 2] = total;
            halo_ms[run / 2] = training ? 0.f : halo;
        }
    }
    return 0;
}

}  // extern "C"

namespace {

// H2D of the host inputs of one step on `stream`.  Default: the whole padded field (the caller's halos are
// periodic).  Interior mode: only the X*Y*Z interior crosses the bus (a pitched 3-D copy) and the periodic
// halo is rebuilt on the device -- the same field as long as the caller's halos were periodic, 8 % fewer bytes.
int upload_inputs(absinthe_variant* v, const double* const* h_inputs, cudaStream_t stream) {
    const Geometry& g = v->geo;
    const size_t bytes = (size_t)g.cells() * sizeof(double);
    FieldList rebuilt{};
    for (size_t n = 0; n < v->inputs.size(); ++n) {
        if (!h_inputs[n]) continue;   // a null entry keeps the device-resident field (constants uploaded once)
        double* dst = v->fields[v->inputs[n]].ptr;
        if (!v->host_interior) {
            RT(cudaMemcpyAsync(dst, h_inputs[n], bytes, cudaMemcpyHostToDevice, stream));
            continue;
        }
        cudaMemcpy3DParms p{};
        const size_t pitch = (size_t)g.row() * sizeof(double);
        p.srcPtr = make_cudaPitchedPtr(const_cast<double*>(h_inputs[n]), pitch, pitch, (size_t)g.rows());
        p.dstPtr = make_cudaPitchedPtr(dst, pitch, pitch, (size_t)g.rows());
        p.srcPos = p.dstPos = make_cudaPos((size_t)g.HX * sizeof(double), (size_t)g.HY, (size_t)g.HZ);
        p.extent = make_cudaExtent((size_t)g.X * sizeof(double), (size_t)g.Y, (size_t)g.Z);
        p.kind = cudaMemcpyHostToDevice;
        RT(cudaMemcpy3DAsync(&p, stream));
        rebuilt.ptr[rebuilt.count++] = dst;
        if (rebuilt.count == kMaxHaloFields) {
            int rc = periodic_update(rebuilt, g, 7, stream);
            if (rc) return rc;
            rebuilt.count = 0;
        }
    }
    return periodic_update(rebuilt, g, 7, stream);
}

}  // namespace

extern "C" {

int absinthe_variant_host_interior(absinthe_variant* v, int enable) {
    if (!v) return fail(-3, "null variant");
    v->host_interior = enable != 0;
    return 0;
}

int absinthe_variant_wait_host(absinthe_variant* v) {
    if (!v) return fail(-3, "null variant");
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d of variant '%s'", v->device, v->name.c_str());
    if (v->host_waited < v->host_issued) {
        RT(cudaEventSynchronize(v->ev_out[v->host_waited % kHostDepth]));
        ++v->host_waited;
    }
    return 0;
}

int absinthe_variant_apply_host(absinthe_variant* v, const double* const* h_inputs,
                                double* const* h_outputs, void* stream) {
    int rc = require_bound(v);
    if (rc) return rc;
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d of variant '%s'", v->device, v->name.c_str());
    while (v->host_waited < v->host_issued)   // async steps still in flight use the same device fields
        if ((rc = absinthe_variant_wait_host(v))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = absinthe_variant_field_bytes(v);
    if ((rc = upload_inputs(v, h_inputs, s))) return rc;
    if ((rc = absinthe_variant_apply(v, s))) return rc;
    for (size_t n = 0; n < v->outputs.size(); ++n)
        RT(cudaMemcpyAsync(h_outputs[n], v->fields[v->outputs[n]].ptr, bytes, cudaMemcpyDeviceToHost, s));
    RT(cudaStreamSynchronize(s));
    return 0;
}

int absinthe_variant_apply_host_async(absinthe_variant* v, const double* const* h_inputs,
                                      double* const* h_outputs, void* stream) {
    int rc = require_bound(v);
    if (rc) return rc;
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d of variant '%s'", v->device, v->name.c_str());
    if (v->host_issued - v->host_waited >= kHostDepth)
        return fail(-6, "%d host steps already in flight: call absinthe_variant_wait_host first", kHostDepth);
    // the copies of this path overlap with the caller and with each other only from page-locked memory; a
    // pageable pointer would silently serialise them (and the call would return before the data is safe)
    for (size_t n = 0; n < v->inputs.size() + v->outputs.size(); ++n) {
        const void* p = n < v->inputs.size() ? (const void*)h_inputs[n] : (const void*)h_outputs[n - v->inputs.size()];
        if (!p) {
            if (n < v->inputs.size()) continue;          // a null input keeps the device-resident field
            return fail(-3, "host output %zu is null", n - v->inputs.size());
        }
        cudaPointerAttributes attr{};
        cudaError_t pe = cudaPointerGetAttributes(&attr, p);
        if (pe != cudaSuccess) { cudaGetLastError(); return fail((int)pe, "cudaPointerGetAttributes: %s", cudaGetErrorString(pe)); }
        if (attr.type != cudaMemoryTypeHost)
            return fail(-7, "absinthe_variant_apply_host_async needs page-locked host buffers (%s %zu is %s memory)",
                        n < v->inputs.size() ? "input" : "output", n < v->inputs.size() ? n : n - v->inputs.size(),
                        attr.type == cudaMemoryTypeUnregistered ? "pageable" : "device");
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (!v->h2d_stream) {
        RT(cudaStreamCreateWithFlags(&v->h2d_stream, cudaStreamNonBlocking));
        RT(cudaStreamCreateWithFlags(&v->d2h_stream, cudaStreamNonBlocking));
        RT(cudaEventCreateWithFlags(&v->ev_in, cudaEventDisableTiming));
        RT(cudaEventCreateWithFlags(&v->ev_done, cudaEventDisableTiming));
        for (cudaEvent_t& e : v->ev_out) RT(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const bool chained = v->host_issued > 0;
    const size_t bytes = absinthe_variant_field_bytes(v);
    // the upload may overwrite the device inputs as soon as the previous step's kernels have read them
    // (ev_done still holds that step's record), i.e. while its outputs are still on their way to the host
    if (chained) RT(cudaStreamWaitEvent(v->h2d_stream, v->ev_done, 0));
    if ((rc = upload_inputs(v, h_inputs, v->h2d_stream))) return rc;
    RT(cudaEventRecord(v->ev_in, v->h2d_stream));
    RT(cudaStreamWaitEvent(s, v->ev_in, 0));
    // ... and the kernels may overwrite the device outputs once the previous step's download has read them
    if (chained) RT(cudaStreamWaitEvent(s, v->ev_out[(v->host_issued - 1) % kHostDepth], 0));
    if ((rc = absinthe_variant_apply(v, s))) return rc;
    RT(cudaEventRecord(v->ev_done, s));
    RT(cudaStreamWaitEvent(v->d2h_stream, v->ev_done, 0));
    for (size_t n = 0; n < v->outputs.size(); ++n)
        RT(cudaMemcpyAsync(h_outputs[n], v->fields[v->outputs[n]].ptr, bytes, cudaMemcpyDeviceToHost, v->d2h_stream));
    RT(cudaEventRecord(v->ev_out[v->host_issued % kHostDepth], v->d2h_stream));
    ++v->host_issued;
    return 0;
}

int absinthe_make_periodic(double* d_field, int X, int Y, int Z, int HX, int HY, int HZ, void* stream) {
    if (!d_field) return fail(-3, "absinthe_make_periodic: null field");
    FieldList list{};
    list.count = 1;
    list.ptr[0] = d_field;
    return periodic_update(list, Geometry{X, Y, Z, HX, HY, HZ}, 7, (cudaStream_t)stream);
}

int absinthe_make_periodic_ik(double* d_field, int X, int Y, int Z, int HX, int HY, int HZ, void* stream) {
    if (!d_field) return fail(-3, "absinthe_make_periodic_ik: null field");
    Geometry g{X, Y, Z, HX, HY, HZ};
    if (!(g.X >= g.HX && g.Z >= g.HZ)) return fail(-3, "absinthe_make_periodic_ik: extent smaller than halo");
    FieldList list{};
    list.count = 1;
    list.ptr[0] = d_field;
    // i faces on interior rows, then k faces on full planes; j-halo rows are refreshed by the
    // neighbour exchange afterwards (which carries their i and k halos with it)
    long total = 2L * g.HZ * g.plane() + (long)g.Z * g.Y * 2 * g.HX;
    periodic_kernel<<<launch_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(list, g, 5);
    RT(cudaGetLastError());
    return 0;
}

void absinthe_fill_rand(double* h_dst, size_t count, unsigned seed, int reseed) {
    if (reseed) std::srand(seed);
    for (size_t n = 0; n < count; ++n) h_dst[n] = std::rand();
}

int absinthe_halo_pack(const double* d_field, double* d_buffer, int X, int Y, int Z, int HX, int HY,
                       int HZ, int j0, int width, void* stream) {
    Geometry g{X, Y, Z, HX, HY, HZ};
    long total = g.planes() * width * g.row();
    pack_kernel<<<launch_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(d_field, d_buffer, g, j0, width, 0);
    RT(cudaGetLastError());
    return 0;
}

int absinthe_halo_unpack(double* d_field, const double* d_buffer, int X, int Y, int Z, int HX, int HY,
                         int HZ, int j0, int width, void* stream) {
    Geometry g{X, Y, Z, HX, HY, HZ};
    long total = g.planes() * width * g.row();
    pack_kernel<<<launch_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(d_field, const_cast<double*>(d_buffer), g, j0, width, 1);
    RT(cudaGetLastError());
    return 0;
}

int absinthe_compare_fields(const double* d_expected, const double* d_actual, int X, int Y, int Z, int HX,
                            int HY, int HZ, double tolerance, unsigned long long* errors,
                            unsigned long long* matches, void* stream) {
    if (!d_expected || !d_actual || !errors || !matches) return fail(-3, "absinthe_compare_fields: null argument");
    Geometry g{X, Y, Z, HX, HY, HZ};
    unsigned long long* counters = nullptr;
    int rc = device_scratch(g_counters, 2 * sizeof(unsigned long long), &counters);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned long long), s);
    if (e == cudaSuccess) {
        compare_kernel<<<launch_grid((long)X * Y * Z, 256), 256, 0, s>>>(d_expected, d_actual, g, tolerance, counters);
        e = cudaGetLastError();
    }
    unsigned long long host[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(host, counters, sizeof(host), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail((int)e, "absinthe_compare_fields: %s", cudaGetErrorString(e));
    *errors = host[0];
    *matches = host[1];
    return 0;
}

}  // extern "C"

// ---- multi-GPU halo exchange (see include/absinthe_b200.h) ---------------------------------------
extern "C" {

int absinthe_comm_create(absinthe_variant* v, int rank, int nranks, unsigned char handle[ABSINTHE_COMM_HANDLE_BYTES]) {
    if (!v || !handle) return fail(-3, "absinthe_comm_create: null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(-3, "absinthe_comm_create: rank %d of %d", rank, nranks);
    static_assert(sizeof(cudaIpcMemHandle_t) == ABSINTHE_COMM_HANDLE_BYTES, "IPC handle size");
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d", v->device);
    absinthe_comm_destroy(v);
    const Geometry& g = v->geo;
    if (g.Y < g.HY) return fail(-3, "slab of %d rows is thinner than the halo", g.Y);
    size_t cursor = 0;
    for (const Group& grp : v->groups) {
        if (grp.halo_fields.empty()) continue;
        absinthe_variant::CommGroup c;
        c.group = grp.id;
        c.face_doubles = grp.halo_fields.size() * (size_t)g.planes() * g.HY * g.row();
        c.flags = cursor;
        cursor += 16;                                     // a 128-byte line for the two arrival epochs
        c.data = cursor;
        cursor += 4 * ((c.face_doubles + 15) / 16 * 16);  // [parity][from_prev | from_next]
        v->comm_groups.push_back(c);
    }
    v->window_doubles = std::max<size_t>(cursor, 16);
    RT(cudaMalloc(&v->window, v->window_doubles * sizeof(double)));
    RT(cudaMemset(v->window, 0, v->window_doubles * sizeof(double)));
    RT(cudaMalloc(&v->put_done, sizeof(unsigned) * std::max<size_t>(v->comm_groups.size(), 1)));
    RT(cudaMemset(v->put_done, 0, sizeof(unsigned) * std::max<size_t>(v->comm_groups.size(), 1)));
    for (auto& c : v->comm_groups) RT(cudaEventCreateWithFlags(&c.unpacked, cudaEventDisableTiming));
    RT(cudaStreamCreateWithFlags(&v->unpack_stream, cudaStreamNonBlocking));
    RT(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    RT(cudaIpcGetMemHandle(&h, v->window));
    std::memcpy(handle, &h, sizeof(h));
    v->comm_rank = rank;
    v->comm_ranks = nranks;
    return 0;
}

int absinthe_comm_connect(absinthe_variant* v, const unsigned char* prev_handle, const unsigned char* next_handle) {
    if (!v || !v->window) return fail(-3, "absinthe_comm_connect: no window (call absinthe_comm_create first)");
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d", v->device);
    if (v->comm_ranks == 1) {                             // both neighbours are this rank
        v->window_prev = v->window_next = v->window;
        return 0;
    }
    if (!prev_handle || !next_handle) return fail(-3, "absinthe_comm_connect: null handle");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, prev_handle, sizeof(h));
    void* mapped = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail((int)e, "cannot map the window of rank %d: %s", (v->comm_rank + v->comm_ranks - 1) % v->comm_ranks, cudaGetErrorString(e));
    v->window_prev = (double*)mapped;
    v->prev_mapped = true;
    if (v->comm_ranks == 2) {                             // the same neighbour on both sides
        v->window_next = v->window_prev;
        return 0;
    }
    std::memcpy(&h, next_handle, sizeof(h));
    e = cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail((int)e, "cannot map the window of rank %d: %s", (v->comm_rank + 1) % v->comm_ranks, cudaGetErrorString(e));
    v->window_next = (double*)mapped;
    v->next_mapped = true;
    return 0;
}

int absinthe_comm_destroy(absinthe_variant* v) {
    if (!v) return 0;
    if (v->window) cudaDeviceSynchronize();
    if (v->prev_mapped && v->window_prev) cudaIpcCloseMemHandle(v->window_prev);
    if (v->next_mapped && v->window_next) cudaIpcCloseMemHandle(v->window_next);
    for (auto& c : v->comm_groups)
        if (c.unpacked) cudaEventDestroy(c.unpacked);
    if (v->unpack_stream) cudaStreamDestroy(v->unpack_stream);
    v->unpack_stream = nullptr;
    if (v->window) cudaFree(v->window);
    if (v->put_done) cudaFree(v->put_done);
    v->window = v->window_prev = v->window_next = nullptr;
    v->put_done = nullptr;
    v->prev_mapped = v->next_mapped = false;
    v->comm_groups.clear();
    v->comm_ranks = 0;
    return 0;
}

static int comm_group(absinthe_variant* v, int group, absinthe_variant::CommGroup** out, Group** grp, size_t* index) {
    int rc = require_bound(v);
    if (rc) return rc;
    if (!v->window_prev || !v->window_next) return fail(-3, "variant '%s': exchange not connected", v->name.c_str());
    for (size_t n = 0; n < v->comm_groups.size(); ++n)
        if (v->comm_groups[n].group == group) {
            *out = &v->comm_groups[n];
            *index = n;
            for (Group& g : v->groups)
                if (g.id == group) *grp = &g;
            return 0;
        }
    *out = nullptr;
    return 0;                                             // a group without outputs has nothing to exchange
}

// strided rows <-> dense buffer with the copy engines: per field one 2-D copy (HY contiguous rows per plane)
static int copy_rows(absinthe_variant* v, const Group& g, double* buffer, long j0, bool to_buffer, cudaStream_t s) {
    const Geometry& q = v->geo;
    const size_t width = (size_t)q.HY * q.row() * sizeof(double);
    for (size_t n = 0; n < g.halo_fields.size(); ++n) {
        double* field = v->fields[g.halo_fields[n]].ptr + j0 * q.row();
        double* dense = buffer + n * (size_t)q.planes() * q.HY * q.row();
        if (to_buffer)
            RT(cudaMemcpy2DAsync(dense, width, field, (size_t)q.plane() * sizeof(double), width, (size_t)q.planes(),
                                 cudaMemcpyDefault, s));
        else
            RT(cudaMemcpy2DAsync(field, (size_t)q.plane() * sizeof(double), dense, width, width, (size_t)q.planes(),
                                 cudaMemcpyDefault, s));
    }
    return 0;
}

static int unpack_on(absinthe_variant* v, absinthe_variant::CommGroup* c, Group* g, cudaStream_t u);

int absinthe_variant_put_group(absinthe_variant* v, int group, int engine, void* stream) {
    absinthe_variant::CommGroup* c = nullptr;
    Group* g = nullptr;
    size_t index = 0;
    int rc = comm_group(v, group, &c, &g, &index);
    if (rc || !c) return rc;
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d", v->device);
    if (engine != ABSINTHE_PUT_KERNEL && engine != ABSINTHE_PUT_COPY) return fail(-3, "unknown PUT engine %d", engine);
    if (c->put != c->waited)
        return fail(-6, "group %d: the previous PUT has not been waited for (absinthe_variant_wait_group)", group);
    cudaStream_t s = (cudaStream_t)stream;
    FieldList list{};
    list.count = (int)g->halo_fields.size();
    for (int n = 0; n < list.count; ++n) list.ptr[n] = v->fields[g->halo_fields[n]].ptr;
    const Geometry& q = v->geo;
    if (!g->fused_halo) {                                 // local periodic i and k faces first: they travel with the rows
        if (!(q.X >= q.HX && q.Z >= q.HZ)) return fail(-3, "extent smaller than halo");
        long cells = 2L * q.HZ * q.plane() + (long)q.Z * q.Y * 2 * q.HX;
        periodic_kernel<<<launch_grid(cells, 256), 256, 0, s>>>(list, q, 5);
    }
    // the neighbours may only overwrite the buffers of this parity once this rank has unpacked them
    if (c->waited >= 1) RT(cudaStreamWaitEvent(s, c->unpacked, 0));
    const unsigned long long epoch = ++c->put;
    c->engine = engine;
    const size_t slot = (c->face_doubles + 15) / 16 * 16;
    const size_t parity = (size_t)(epoch & 1);
    // my low rows -> previous rank's "from next" buffer; my high rows -> next rank's "from previous" buffer
    double* to_prev = v->window_prev + c->data + (2 * parity + 1) * slot;
    double* to_next = v->window_next + c->data + (2 * parity + 0) * slot;
    unsigned long long* flag_at_prev = (unsigned long long*)(v->window_prev + c->flags) + 1;
    unsigned long long* flag_at_next = (unsigned long long*)(v->window_next + c->flags) + 0;
    if (engine == ABSINTHE_PUT_KERNEL) {
        const int lines = list.count * (int)q.planes() * q.HY;
        put_kernel<<<std::min(lines, sm_count()), kExchangeThreads, 0, s>>>(list, q, to_prev, to_next, flag_at_prev,
                                                                            flag_at_next, v->put_done + index, epoch);
        RT(cudaGetLastError());
        return 0;
    }
    // the unpack of this exchange goes to the library's stream right away: behind the group's kernel (everything
    // `s` holds up to here), it waits for the neighbours' flags and fills the j-halos while the caller's stream
    // computes the next groups; the WAIT then only waits for its event
    RT(cudaEventRecord(c->unpacked, s));
    RT(cudaStreamWaitEvent(v->unpack_stream, c->unpacked, 0));
    if ((rc = copy_rows(v, *g, to_prev, q.HY, true, s))) return rc;
    if ((rc = copy_rows(v, *g, to_next, q.Y, true, s))) return rc;
    bool raised = false;
    if (g_drv.StreamWriteValue64) {                       // ordered after the copies, with a memory barrier
        CUresult r1 = g_drv.StreamWriteValue64((CUstream)s, (CUdeviceptr)flag_at_prev, epoch, CU_STREAM_WRITE_VALUE_DEFAULT);
        CUresult r2 = g_drv.StreamWriteValue64((CUstream)s, (CUdeviceptr)flag_at_next, epoch, CU_STREAM_WRITE_VALUE_DEFAULT);
        raised = r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS;
        if (!raised) g_drv.StreamWriteValue64 = nullptr;  // not on this memory / driver: raise the flags from a kernel
    }
    if (!raised) {
        signal_kernel<<<1, 1, 0, s>>>(flag_at_prev, flag_at_next, epoch);
        RT(cudaGetLastError());
    }
    return unpack_on(v, c, g, v->unpack_stream);
}

int absinthe_variant_apply_group_put(absinthe_variant* v, int group, void* stream) {
    absinthe_variant::CommGroup* c = nullptr;
    Group* g = nullptr;
    size_t index = 0;
    int rc = comm_group(v, group, &c, &g, &index);
    if (rc) return rc;
    if (!c || !g->fused_halo) {                           // nothing to push from inside the kernel: kernel, then PUT
        if ((rc = absinthe_variant_apply_group(v, group, 0, stream))) return rc;
        return c ? absinthe_variant_put_group(v, group, ABSINTHE_PUT_KERNEL, stream) : 0;
    }
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d", v->device);
    if (c->put != c->waited)
        return fail(-6, "group %d: the previous PUT has not been waited for (absinthe_variant_wait_group)", group);
    cudaStream_t s = (cudaStream_t)stream;
    if (c->waited >= 1) RT(cudaStreamWaitEvent(s, c->unpacked, 0));
    const unsigned long long epoch = c->put + 1;
    const size_t slot = (c->face_doubles + 15) / 16 * 16;
    const size_t parity = (size_t)(epoch & 1);
    double* to_prev = v->window_prev + c->data + (2 * parity + 1) * slot;
    double* to_next = v->window_next + c->data + (2 * parity + 0) * slot;
    if (!set_remote(*g, to_prev, to_next)) return fail(-3, "kernel of group %d has no slot for the exchange windows", group);
    rc = launch_group(v, *g, false, s);
    set_remote(*g, nullptr, nullptr);                     // plain applications keep writing the rank's own halo
    if (rc) return rc;
    c->put = epoch;
    c->engine = ABSINTHE_PUT_KERNEL;
    unsigned long long* flag_at_prev = (unsigned long long*)(v->window_prev + c->flags) + 1;
    unsigned long long* flag_at_next = (unsigned long long*)(v->window_next + c->flags) + 0;
    if (g_drv.StreamWriteValue64) {                       // after the kernel, with a memory barrier
        CUresult r1 = g_drv.StreamWriteValue64((CUstream)s, (CUdeviceptr)flag_at_prev, epoch, CU_STREAM_WRITE_VALUE_DEFAULT);
        CUresult r2 = g_drv.StreamWriteValue64((CUstream)s, (CUdeviceptr)flag_at_next, epoch, CU_STREAM_WRITE_VALUE_DEFAULT);
        if (r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS) return 0;
        g_drv.StreamWriteValue64 = nullptr;
    }
    signal_kernel<<<1, 1, 0, s>>>(flag_at_prev, flag_at_next, epoch);
    RT(cudaGetLastError());
    return 0;
}

// wait for both neighbours' flags of the outstanding PUT on stream `u` and fill the j-halos; records `unpacked`
static int unpack_on(absinthe_variant* v, absinthe_variant::CommGroup* c, Group* g, cudaStream_t u) {
    FieldList list{};
    list.count = (int)g->halo_fields.size();
    for (int n = 0; n < list.count; ++n) list.ptr[n] = v->fields[g->halo_fields[n]].ptr;
    const Geometry& q = v->geo;
    const unsigned long long epoch = c->put;
    const size_t slot = (c->face_doubles + 15) / 16 * 16;
    const size_t parity = (size_t)(epoch & 1);
    double* from_prev = v->window + c->data + (2 * parity + 0) * slot;
    double* from_next = v->window + c->data + (2 * parity + 1) * slot;
    unsigned long long* flags = (unsigned long long*)(v->window + c->flags);
    bool waited = false;
    if (g_drv.StreamWaitValue64) {
        CUresult r1 = g_drv.StreamWaitValue64((CUstream)u, (CUdeviceptr)flags, epoch, CU_STREAM_WAIT_VALUE_GEQ);
        CUresult r2 = g_drv.StreamWaitValue64((CUstream)u, (CUdeviceptr)(flags + 1), epoch, CU_STREAM_WAIT_VALUE_GEQ);
        waited = r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS;
        if (!waited) g_drv.StreamWaitValue64 = nullptr;
    }
    if (!waited) {
        await_kernel<<<1, 1, 0, u>>>(flags, epoch);
        RT(cudaGetLastError());
    }
    if (c->engine == ABSINTHE_PUT_KERNEL) {
        const int lines = list.count * (int)q.planes() * q.HY;
        unpack_kernel<<<std::min(lines, sm_count()), kExchangeThreads, 0, u>>>(list, q, from_prev, from_next);
        RT(cudaGetLastError());
    } else {
        int rc = 0;
        if ((rc = copy_rows(v, *g, from_prev, 0, false, u))) return rc;
        if ((rc = copy_rows(v, *g, from_next, q.Y + q.HY, false, u))) return rc;
    }
    RT(cudaEventRecord(c->unpacked, u));
    return 0;
}

int absinthe_variant_wait_group(absinthe_variant* v, int group, void* stream) {
    absinthe_variant::CommGroup* c = nullptr;
    Group* g = nullptr;
    size_t index = 0;
    int rc = comm_group(v, group, &c, &g, &index);
    if (rc || !c) return rc;
    if (c->waited == c->put) return 0;                    // nothing outstanding
    OnDevice on(v->device);
    if (!on.ok) return fail(-3, "cannot switch to device %d", v->device);
    if (c->engine == ABSINTHE_PUT_KERNEL) {
        // halos the very next kernel reads: wait and unpack on the caller's own stream, no event hops
        if ((rc = unpack_on(v, c, g, (cudaStream_t)stream))) return rc;
    } else {
        // copy engine: the unpack was enqueued with the PUT (on the library's stream, behind the group's kernel)
        // and has usually long finished; the caller's stream only waits for its event
        RT(cudaStreamWaitEvent((cudaStream_t)stream, c->unpacked, 0));
    }
    c->waited = c->put;
    return 0;
}

}  // extern "C"
