// sgm_b200.cu -- host side and C-ABI of libsgm_b200.so (see include/SemiGlobalMatching.h, include/sgm_b200.h).
//
// Pipeline per frame (all on one stream, no host round trip in between):
//   memset side buffer -> K1 census (both images) -> K2 aggregation (all paths of all directions, one
//   launch) -> K3 plane sum + WTA(left,right) + sub-pixel + LR check -> [K4 speckle CCL] -> [K5 in-place
//   median wavefront].
// There is no CPU implementation behind these entry points: without a CUDA device they fail.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sgm_b200.h"
#include "aggregate.cuh"
#include "census.cuh"
#include "frontend.cuh"
#include "path_walker.h"
#include "postproc.cuh"
#include "wta.cuh"

using namespace sgmb;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(SGMB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* SGMB_LastError(void) { return g_err; }

// ------------------------------------------------------------------------------------------------ context
constexpr int kMaxTimedKernels = 16;
struct KernelTimer {                  // one event before the first launch, one after every launch
    cudaEvent_t ev[kMaxTimedKernels + 1] = {};
    const char* name[kMaxTimedKernels] = {};
    int n = 0;
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr, evAgg0 = nullptr, evAgg1 = nullptr, evDone = nullptr, evCopied = nullptr;
    uint8_t* img[2] = {nullptr, nullptr};
    void* censusL = nullptr;          // descriptors: uint32 (5x5 census) or 64-bit (9x7 census)
    void* censusR4 = nullptr;         // [32 / descBytes][copyStride] shifted copies (census.cuh)
    void* pixL = nullptr;             // {left descriptor, grey} per pixel: uint2 / uint4 for 32- / 64-bit descriptors (aggregate.cuh)
    uint8_t* planes = nullptr;
    uint16_t* side = nullptr;
    uint16_t* S = nullptr;            // taps only
    float* dispLeftWta = nullptr;     // taps only
    float* dispRight = nullptr;       // taps only
    float* dispLR = nullptr;
    float* rightRow = nullptr;        // K3 scratch: right-view disparities
    uint4* wtaRecords = nullptr;      // K3 scratch: [2][N] parked scan results
    int* wtaRowDone = nullptr;        // K3: arrival counters of rows split over several blocks (zero between launches)
    int* wtaSched = nullptr;          // K3: per-SM block counters + work tickets (zero between launches)
    float* dispSpeckle = nullptr;
    float* dispFinal = nullptr;
    int32_t* labels = nullptr;        // speckle filter scratch [2N]
    unsigned long long* xchg = nullptr;   // median wavefront exchange rows [(H+31)/32][W]
    float* medianPrep = nullptr;      // sorted unfiltered inputs in wavefront order (postproc.cuh K5a)
    uint8_t* framePlanes = nullptr;   // SGMB_MatchFrame: left B,G,R then right B,G,R planes [6][N] (allocated on first use)
    float* depth = nullptr;           // SGMB_MatchFrame with calibration: depth map [N] (allocated on first use)
    unsigned medianEpoch = 0;
    // SGMB_Match / SGMB_MatchBatch always run the frame on the slot's own image buffers, so the whole frame (memsets and
    // kernels) is recorded once into a CUDA graph and replayed by every later call until the configuration changes
    cudaGraphExec_t frameExec = nullptr;
    // the recorded graph itself (its node handles are needed to re-point the last kernel at the caller's buffer), the median
    // wavefront's node in it, and the output pointer the executable graph currently holds
    cudaGraph_t frameGraph = nullptr;
    cudaGraphNode_t medianNode = nullptr;
    float* frameOut = nullptr;
    bool busy = false;
    // Pageable caller memory (malloc / static arrays, e.g. the reference demo main.c:25-26,81) is staged through page-locked
    // buffers owned by the slot (allocated on first use): stageIn = both images in the layout of img[], stageOut = result.
    uint8_t* stageIn = nullptr;
    float* stageOut = nullptr;
    float* pendingOut = nullptr;      // pageable destination that stageOut still has to be copied to (after the stream drained)
    KernelTimer* timer = nullptr;     // SGMB_TimeKernels: events around every launch of enqueue_frame
};

struct SGMB_Context {
    int device = 0;
    int nslots = 1;
    std::vector<Slot> slots;
    bool configured = false;
    unsigned pipeline = SGMB_PIPE_REFERENCE;
    SGMOption opt{};
    int W = 0, H = 0, D = 0, Dp = 0, NR = 0, nDirs = 8;
    int censusW = 5, censusH = 5;     // requested window (SGMB_SetCensusWindow); takes effect at SGMB_Configure
    int descBytes = 4;                // descriptor size of the current configuration: 4 (5x5) or 8 (9x7)
    int greyFormula = SGMB_GREY_BOARD; // SGMB_MatchFrame: weights of the colour -> grey conversion
    CompareAcc* cmpScratch = nullptr; // SGMB_CompareDepth*: per-block partials + result
    size_t N = 0;
    size_t imgStride = 0;             // bytes between the left and the right image of a slot (N + 16 rounded up to 256)
    int padF = 0;
    size_t copyStride = 0, planeStride = 0;
    // configuration-wide read-only device tables
    WarpWork* work = nullptr;
    int nIrregularWarps = 0, nRegularWarps = 0;
    int lppV = 8;                 // lanes per path of the vertical / diagonal directions
    int lppH = 16;                // lanes per path of the horizontal directions
    int altLayout = 0;            // SGM_B200_DEBUG_LAYOUT: alternative kernel layouts for experiments
    int32_t* entryOf = nullptr;
    int nEntries = 0, nIrregular = 0;
    uint32_t p2x2[256] = {};
    uint32_t p1x2 = 0;
    // K3 launch shape: lanes per pixel (power of two >= Dp / 16), columns per tile, dynamic shared memory
    int wtaCPP = 8, wtaTW = 32, wtaRingRows = 0;
    size_t wtaSmem = 0;
    int smCount = 148;
    int wtaRowsPerSm = 0;
    int wtaFullRows = 0, wtaPieces = 0;   // K3 work split (wta_plan): whole-row blocks + pieces of the last partial wave of rows
    int4* wtaSegments = nullptr;
    // L2 flush scratch for SGMB_TimeDevice
    uint8_t* flushBuf = nullptr;
    size_t flushBytes = 0;
    float lastMs = 0.f;
    bool tapsAllocated = false;
    bool capturing = false;       // enqueue_frame is being recorded into a CUDA graph (SGMB_RunDevice)
    const char* timedName[kMaxTimedKernels] = {};   // kernel names of the last SGMB_TimeKernels call
};

static int ensure_device(SGMB_Context* c)
{
    CU(cudaSetDevice(c->device));
    return SGMB_OK;
}

static void drop_frame_graph(Slot& s)
{
    if (s.frameExec) { cudaGraphExecDestroy(s.frameExec); s.frameExec = nullptr; }
    if (s.frameGraph) { cudaGraphDestroy(s.frameGraph); s.frameGraph = nullptr; }
    s.medianNode = nullptr; s.frameOut = nullptr;
}

static void free_slot_buffers(Slot& s)
{
    drop_frame_graph(s);
    cudaFree(s.img[0]); cudaFree(s.censusL);
    if (s.stageIn) cudaFreeHost(s.stageIn);
    if (s.stageOut) cudaFreeHost(s.stageOut);
    s.stageIn = nullptr; s.stageOut = nullptr; s.pendingOut = nullptr; cudaFree(s.censusR4); cudaFree(s.planes);
    cudaFree(s.side); cudaFree(s.S); cudaFree(s.dispLeftWta); cudaFree(s.dispRight); cudaFree(s.dispLR);
    cudaFree(s.framePlanes); cudaFree(s.depth); s.framePlanes = nullptr; s.depth = nullptr;
    cudaFree(s.pixL); s.pixL = nullptr;
    cudaFree(s.dispSpeckle); cudaFree(s.dispFinal); cudaFree(s.labels); cudaFree(s.xchg); cudaFree(s.medianPrep); cudaFree(s.rightRow); cudaFree(s.wtaRecords); cudaFree(s.wtaRowDone); cudaFree(s.wtaSched);
    s.img[0] = s.img[1] = nullptr; s.censusL = s.censusR4 = nullptr; s.planes = nullptr; s.side = nullptr; s.S = nullptr;
    s.dispLeftWta = s.dispRight = s.dispLR = s.dispSpeckle = s.dispFinal = nullptr; s.labels = nullptr; s.xchg = nullptr; s.medianPrep = nullptr; s.rightRow = nullptr; s.wtaRecords = nullptr; s.wtaRowDone = nullptr; s.wtaSched = nullptr;
}

static void free_config(SGMB_Context* c)
{
    for (auto& s : c->slots) free_slot_buffers(s);
    cudaFree(c->work); cudaFree(c->entryOf); cudaFree(c->wtaSegments);
    c->work = nullptr; c->entryOf = nullptr; c->wtaSegments = nullptr;
    c->configured = false; c->tapsAllocated = false;
}

extern "C" int SGMB_DeviceCount(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(SGMB_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

extern "C" int SGMB_Create(SGMB_Context** out, int device, int slots)
{
    if (!out || slots < 1 || slots > 64) return fail(SGMB_E_ARG, "SGMB_Create: bad arguments");
    *out = nullptr;
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(SGMB_E_CUDA, "SGMB_Create: device %d not present (%d visible)", device, n);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(SGMB_E_CUDA, "SGMB_Create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    auto* c = new SGMB_Context();
    c->device = device;
    c->smCount = prop.multiProcessorCount;
    c->nslots = slots;
    c->slots.resize(slots);
    const int rc = [&]() -> int {
        CU(cudaSetDevice(device));
        for (auto& s : c->slots) {
            CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            CU(cudaEventCreate(&s.evStart)); CU(cudaEventCreate(&s.evStop));
            CU(cudaEventCreate(&s.evAgg0)); CU(cudaEventCreate(&s.evAgg1));
            CU(cudaEventCreateWithFlags(&s.evDone, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s.evCopied, cudaEventDisableTiming));
        }
        return SGMB_OK;
    }();
    if (rc != SGMB_OK) { SGMB_Destroy(c); return rc; }      // releases whatever was created (g_err keeps the reason)
    *out = c;
    return SGMB_OK;
}

extern "C" void SGMB_Destroy(SGMB_Context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_config(c);
    cudaFree(c->flushBuf);
    cudaFree(c->cmpScratch);
    for (auto& s : c->slots) {
        if (s.evStart) cudaEventDestroy(s.evStart);
        if (s.evStop) cudaEventDestroy(s.evStop);
        if (s.evAgg0) cudaEventDestroy(s.evAgg0);
        if (s.evAgg1) cudaEventDestroy(s.evAgg1);
        if (s.evDone) cudaEventDestroy(s.evDone);
        if (s.evCopied) cudaEventDestroy(s.evCopied);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete c;
}

// ------------------------------------------------------------------------------------------------ path topology (host)
// A path is regular iff its walk equals the toroidal diagonal at every visit (path_walker.h).
static bool path_is_regular(int W, int H, int d, int path)
{
    if (d < 4) return true;
    const Dir dir = direction(d);
    PathWalker wk;
    wk.start(W, H, dir.dx, dir.dy, path);
    for (int s = 0; s < H; ++s) {
        if (s) wk.advance();
        if (wk.pos != regular_position(W, H, dir.dx, dir.dy, path, s)) return false;
    }
    return true;
}

extern "C" int SGMB_DebugWalkPath(int W, int H, int d, int path, int* out, int capacity)
{
    if (W <= 0 || H <= 0 || d < 0 || d > 7 || !out) return fail(SGMB_E_ARG, "SGMB_DebugWalkPath: bad arguments");
    const Dir dir = direction(d);
    PathWalker wk;
    wk.start(W, H, dir.dx, dir.dy, path);
    const int len = wk.length();
    if (path < 0 || path >= (d < 2 ? H : W) || capacity < len) return fail(SGMB_E_ARG, "SGMB_DebugWalkPath: bad path/capacity");
    for (int s = 0; s < len; ++s) {
        if (s) wk.advance();
        out[s] = wk.pos;
    }
    return len;
}

extern "C" int SGMB_DebugClassifyPaths(int W, int H, int d, uint8_t* irregular, int capacity)
{
    if (W <= 0 || H <= 0 || d < 0 || d > 7 || !irregular) return fail(SGMB_E_ARG, "SGMB_DebugClassifyPaths: bad arguments");
    const int n = d < 2 ? H : W;
    if (capacity < n) return fail(SGMB_E_ARG, "SGMB_DebugClassifyPaths: capacity");
    for (int i = 0; i < n; ++i) irregular[i] = path_is_regular(W, H, d, i) ? 0 : 1;
    return n;
}

// ------------------------------------------------------------------------------------------------ configure
static int alloc_taps(SGMB_Context* c)
{
    if (c->tapsAllocated) return SGMB_OK;
    for (auto& s : c->slots) {
        CU(cudaMalloc(&s.S, c->N * (size_t)c->D * sizeof(uint16_t)));
        CU(cudaMalloc(&s.dispLeftWta, c->N * sizeof(float)));
        CU(cudaMalloc(&s.dispRight, c->N * sizeof(float)));
    }
    c->tapsAllocated = true;
    return SGMB_OK;
}

template <int CPP>
static int wta_prepare(SGMB_Context* c)
{
    c->wtaCPP = CPP;
    c->wtaTW = WtaShape<CPP>::kTW;
    c->wtaSmem = WtaShape<CPP>::ring_bytes(c->D);
    c->wtaRingRows = WtaShape<CPP>::ring_rows(c->D);
    // The attribute belongs to the function on this device, i.e. to every context of the process: it is set to the largest
    // ring of the CPP class (D = 16 * CPP), never to this context's own size - a later context with a smaller D must not
    // lower the limit under a live context with a larger one.
    const int classMax = (int)WtaShape<CPP>::ring_bytes(16 * CPP);
    CU(cudaFuncSetAttribute(sgm_reduce_wta_lr<CPP, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, classMax));
    CU(cudaFuncSetAttribute(sgm_reduce_wta_lr<CPP, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, classMax));
    CU(cudaFuncSetAttribute(sgm_reduce_wta_lr<CPP, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, classMax));
    CU(cudaFuncSetAttribute(sgm_reduce_wta_lr<CPP, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, classMax));
    return SGMB_OK;
}

// K3 work split.  One block per row leaves the SMs unevenly loaded when H is a small, non-integral multiple of the SM count
// (375 rows on 148 SMs: 79 SMs sum three rows, 69 two, and the kernel takes as long as 444 rows would).  The rows of the
// last partial wave are therefore cut, as one strip of tiles, into one equal piece per SM; a piece is at most two segments
// (it may cross one row boundary), launched after the whole-row blocks.  A segment pays for the D - 1 halo columns to its
// right (plane sum only), so rows are only split when a piece is at least twice as long as its halo.
static int wta_plan(SGMB_Context* c)
{
    const char* sms = getenv("SGM_B200_DEBUG_WTA_SMS");         // tests: plan as if the device had this many SMs
    const int W = c->W, H = c->H, TW = c->wtaTW, S = std::max(1, sms && atoi(sms) > 0 ? atoi(sms) : c->smCount);
    const int waves = H / S, rem = H % S;
    const int nT = (W + TW - 1) / TW, haloT = (c->D - 1 + TW - 1) / TW;
    c->wtaFullRows = H; c->wtaPieces = 0;
    const char* e = getenv("SGM_B200_DEBUG_NOSPLIT");
    if (rem == 0 || waves >= 8 || (e && atoi(e))) return SGMB_OK;
    const int total = rem * nT, q = (total + S - 1) / S;
    if (q < 2 * haloT || q < 2) return SGMB_OK;
    const int pieces = (total + q - 1) / q, first = H - rem;
    std::vector<int4> seg(2 * (size_t)pieces, make_int4(-1, 0, 0, 0));
    std::vector<int> perRow(rem, 0);
    for (int pass = 0; pass < 2; ++pass)
        for (int p = 0; p < pieces; ++p) {
            const int t0 = p * q, t1 = std::min(total, t0 + q);
            int k = 0;
            for (int r = t0 / nT; r <= (t1 - 1) / nT; ++r, ++k) {
                const int a = std::max(t0, r * nT) - r * nT, b = std::min(t1, (r + 1) * nT) - r * nT;   // tiles [a, b) of strip row r
                if (pass == 0) ++perRow[r];
                else seg[2 * p + k] = make_int4(first + r, a * TW, std::min(W, b * TW), perRow[r]);
            }
        }
    CU(cudaMalloc(&c->wtaSegments, seg.size() * sizeof(int4)));
    CU(cudaMemcpy(c->wtaSegments, seg.data(), seg.size() * sizeof(int4), cudaMemcpyHostToDevice));
    c->wtaFullRows = first; c->wtaPieces = pieces; c->wtaRowsPerSm = waves;
    return SGMB_OK;
}

extern "C" int SGMB_Configure(SGMB_Context* c, uint16_t width, uint16_t height, const SGMOption* option)
{
    if (!c) return fail(SGMB_E_ARG, "SGMB_Configure: NULL context");
    if (!option) return fail(SGMB_E_ARG, "SGMB_Configure: NULL option");
    if (int rc = ensure_device(c)) return rc;
    CU(cudaDeviceSynchronize());
    free_config(c);
    // the reference's own argument checks (SemiGlobalMatching.c:43-48)
    if (width == 0 || height == 0) return fail(SGMB_E_ARG, "width/height must be non-zero");
    if (option->max_disparity <= option->min_disparity) return fail(SGMB_E_ARG, "max_disparity must exceed min_disparity");
    const int D = option->max_disparity - option->min_disparity;
    if (D > 256) return fail(SGMB_E_UNSUPPORTED, "disparity range %d > 256 is not supported", D);
    if (option->p1 < 0 || option->p2_init < 0) return fail(SGMB_E_UNSUPPORTED, "negative P1/P2 are not supported");
    if ((size_t)width * height >= ((size_t)1 << 30)) return fail(SGMB_E_UNSUPPORTED, "image too large");

    c->opt = *option;
    c->W = width; c->H = height; c->D = D;
    c->N = (size_t)width * height;
    c->imgStride = (c->N + 16 + 255) & ~(size_t)255;
    c->NR = D <= 64 ? 1 : (D <= 128 ? 2 : 4);
    c->descBytes = (c->censusW == 9 && c->censusH == 7) ? 8 : 4;
    const int nCopies = 32 / c->descBytes;            // shifted copies of the right census: any window starts 32-byte aligned in one of them
    c->Dp = (D + 15) & ~15;
    c->nDirs = (option->num_paths == 4) ? 4 : 8;      // the reference ignores num_paths (always 8)
    c->padF = ((option->min_disparity + 64 * c->NR + 8) + 7) & ~7;
    c->copyStride = ((size_t)c->padF + c->N + 16 + 7) & ~(size_t)7;
    c->planeStride = c->N * (size_t)c->Dp;
    // K3 addresses the planes with 32-bit indices in 16-byte units (wta.cuh)
    if (7 * (c->planeStride >> 4) >= ((size_t)1 << 32))
        return fail(SGMB_E_UNSUPPORTED, "frame too large: width * height * disparity range must stay below 9.8e9");
    const int W = c->W, H = c->H;

    // ---- K3 launch shape
    {
        const int chunks = c->Dp / 16;
        int rc;
        if (chunks <= 1)      rc = wta_prepare<1>(c);
        else if (chunks <= 2) rc = wta_prepare<2>(c);
        else if (chunks <= 4) rc = wta_prepare<4>(c);
        else if (chunks <= 8) rc = wta_prepare<8>(c);
        else                  rc = wta_prepare<16>(c);
        if (rc == SGMB_OK) rc = wta_plan(c);
        if (rc) return rc;
        CU(median_configure());
        if (rc) return rc;
    }

    // ---- path classification + work list (host walk of the 4*W diagonal paths; same walker as the kernel).
    //      32/lppH rows per warp for the horizontal directions, 32/lppV paths of one direction per warp otherwise.
    //      Horizontal jobs first (measured: any other position is 6-8 % slower); all CTAs are co-resident at C2.
    //      64-bit descriptors double the registers of the prefetched census windows, so that mode keeps 8
    //      disparities per lane up to D = 128 (must match the template arguments in enqueue_frame).
    c->lppV = (D <= 128 && !(c->descBytes == 8 && D > 64)) ? 8 : 16;
    c->altLayout = getenv("SGM_B200_DEBUG_LAYOUT") ? atoi(getenv("SGM_B200_DEBUG_LAYOUT")) : 0;   // experiments only
    if (c->altLayout == 1 && c->descBytes == 4 && c->NR == 2) c->lppV = 16;
    c->lppH = 16;
    const int perWarpV = 32 / c->lppV;
    const int perWarpH = 32 / c->lppH;
    std::vector<WarpWork> irregular, regular;
    std::vector<int32_t> entryOf(c->N, -1);
    int nEntries = 0;
    // Profiling aid only (results are wrong when set): SGM_B200_DEBUG_DIRMASK keeps just the directions in the bit mask.
    const char* dm = getenv("SGM_B200_DEBUG_DIRMASK");
    const unsigned dirMask = dm ? (unsigned)strtoul(dm, nullptr, 0) : 0xFFu;
    auto push_paths = [&](int d, const std::vector<int>& paths, int perWarp) {
        if (!((dirMask >> d) & 1u)) return;
        size_t i = 0;
        while (i < paths.size()) {                 // consecutive path indices only, so a warp's groups are first..first+count-1
            size_t j = i + 1;
            while (j < paths.size() && paths[j] == paths[j - 1] + 1) ++j;     // [i, j): a run of consecutive paths
            const size_t run = j - i;
            for (size_t k = 0; k < run; k += perWarp) {
                size_t n = std::min<size_t>(perWarp, run - k), first = i + k;
                regular.push_back(WarpWork{paths[first], (uint8_t)d, (uint8_t)n, 0});
            }
            i = j;
        }
    };
    auto all_paths = [&](int n) { std::vector<int> v(n); for (int i = 0; i < n; ++i) v[i] = i; return v; };
    // longest paths first: the horizontal directions have only H paths of W steps each
    // Landscape frames: the vertical directions go LAST.  All blocks of a KITTI-sized frame are resident at once, but the warp
    // schedulers serve the warps of earlier blocks first (per-warp timeline, scripts/micro/agg_trace.py: the four blocks of an SM
    // finish ~25 us apart in launch order although they start together), so the block that started last runs the end of its
    // job alone, at a lone warp's issue rate; the vertical jobs are the cheapest per visit (no column wrap, border test once
    // per warp) and make the shortest tail: 264.8 -> 262.7 us at C2.
    const bool verticalLast = W >= H && !getenv("SGM_B200_DEBUG_VFIRST");
    if (W >= H) { push_paths(0, all_paths(H), perWarpH); push_paths(1, all_paths(H), perWarpH); if (!verticalLast) { push_paths(2, all_paths(W), perWarpV); push_paths(3, all_paths(W), perWarpV); } }
    else        { push_paths(2, all_paths(W), perWarpV); push_paths(3, all_paths(W), perWarpV); push_paths(0, all_paths(H), perWarpH); push_paths(1, all_paths(H), perWarpH); }
    for (int d = 4; d < c->nDirs; ++d) {
        const Dir dir = direction(d);
        std::vector<int> reg;
        for (int i = 0; i < W; ++i) {
            if (path_is_regular(W, H, d, i)) { reg.push_back(i); continue; }
            irregular.push_back(WarpWork{i, (uint8_t)d, 1, 0});
            PathWalker wk;
            wk.start(W, H, dir.dx, dir.dy, i);
            for (int s = 0; s < H; ++s) {
                if (s) wk.advance();
                if (wk.inside() && entryOf[wk.pos] < 0) entryOf[wk.pos] = nEntries++;
            }
        }
        push_paths(d, reg, perWarpV);
    }
    if (verticalLast) { push_paths(2, all_paths(W), perWarpV); push_paths(3, all_paths(W), perWarpV); }
    if (getenv("SGM_B200_DEBUG_NOIRR")) irregular.clear();      // profiling aid only (results are wrong when set)
    std::vector<WarpWork> work(irregular);
    work.insert(work.end(), regular.begin(), regular.end());
    c->nIrregularWarps = (int)irregular.size();
    c->nRegularWarps = (int)(work.size() - irregular.size());
    c->nEntries = nEntries;
    c->nIrregular = (int)irregular.size();
    CU(cudaMalloc(&c->work, work.size() * sizeof(WarpWork)));
    CU(cudaMemcpy(c->work, work.data(), work.size() * sizeof(WarpWork), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&c->entryOf, c->N * sizeof(int32_t)));
    CU(cudaMemcpy(c->entryOf, entryOf.data(), c->N * sizeof(int32_t), cudaMemcpyHostToDevice));

    // ---- penalty table: min(256, max(P1, P2_init / (delta + 1)))  (SemiGlobalMatching.c:335) replicated into both
    //      16-bit fields; candidates above 255 can never be selected, so clamping at 256 keeps the fields from
    //      overflowing.  The table travels in the kernel parameters (constant bank).
    for (int dg = 0; dg < 256; ++dg)
        c->p2x2[dg] = (uint32_t)std::min(256, std::max((int)option->p1, (int)option->p2_init / (dg + 1))) * 0x00010001u;
    c->p1x2 = (uint32_t)std::min(256, (int)option->p1) * 0x00010001u;

    // ---- per-slot device buffers
    for (auto& s : c->slots) {
        CU(cudaMalloc(&s.img[0], 2 * c->imgStride));            // left and right image in one allocation: one H2D copy from staging
        s.img[1] = s.img[0] + c->imgStride;
        CU(cudaMalloc(&s.censusL, c->N * (size_t)c->descBytes));
        CU(cudaMalloc(&s.censusR4, nCopies * c->copyStride * (size_t)c->descBytes));
        CU(cudaMemset(s.censusR4, 0, nCopies * c->copyStride * (size_t)c->descBytes));
        CU(cudaMalloc(&s.pixL, c->N * (c->descBytes == 4 ? sizeof(uint2) : sizeof(uint4))));
        CU(cudaMalloc(&s.planes, (size_t)c->nDirs * c->planeStride));
        // slots on an irregular path's toroidal diagonal are never written by K2 and must read as 0 in K3
        CU(cudaMemset(s.planes, 0, (size_t)c->nDirs * c->planeStride));
        CU(cudaMalloc(&s.side, std::max<size_t>(1, (size_t)nEntries) * c->Dp * sizeof(uint16_t) + 64));
        CU(cudaMalloc(&s.dispLR, c->N * sizeof(float)));
        CU(cudaMalloc(&s.rightRow, c->N * sizeof(float)));
        CU(cudaMalloc(&s.wtaRecords, 2 * c->N * sizeof(uint4)));
        CU(cudaMalloc(&s.wtaRowDone, (size_t)H * sizeof(int)));
        CU(cudaMemset(s.wtaRowDone, 0, (size_t)H * sizeof(int)));
        CU(cudaMalloc(&s.wtaSched, kWtaSchedInts * sizeof(int)));
        CU(cudaMemset(s.wtaSched, 0, kWtaSchedInts * sizeof(int)));
        CU(cudaMalloc(&s.dispSpeckle, c->N * sizeof(float)));
        CU(cudaMalloc(&s.dispFinal, c->N * sizeof(float)));
        CU(cudaMalloc(&s.labels, 2 * c->N * sizeof(int32_t)));
        const size_t xbytes = median_xchg_bytes(W, H);
        CU(cudaMalloc(&s.xchg, xbytes));
        CU(cudaMemset(s.xchg, 0, xbytes));
        // slots of idle (row, step) pairs are never written and must hold ordinary floats
        CU(cudaMalloc(&s.medianPrep, median_prep_floats(W, H) * sizeof(float)));
        CU(cudaMemset(s.medianPrep, 0, median_prep_floats(W, H) * sizeof(float)));
        s.medianEpoch = 0;
        s.busy = false;
    }
    if (c->pipeline & SGMB_PIPE_TAPS) if (int rc = alloc_taps(c)) return rc;
    CU(cudaDeviceSynchronize());
    c->configured = true;
    return SGMB_OK;
}

#ifdef SGM_SPECKLE_DEBUG
extern "C" int SGMB_DebugUfStats(unsigned long long* dst, int reset)
{
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(dst, g_ufStats, sizeof(g_ufStats)));
    if (reset) { unsigned long long z[8] = {}; CU(cudaMemcpyToSymbol(g_ufStats, z, sizeof z)); }
    return SGMB_OK;
}
#endif

#ifdef SGM_AGG_TRACE
extern "C" int SGMB_DebugAggTrace(unsigned long long* dst, int nWarps)
{
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(dst, g_aggTrace, (size_t)std::min(nWarps, kAggTraceWarps) * 3 * sizeof(unsigned long long)));
    return SGMB_OK;
}
#endif

extern "C" int SGMB_SetCensusWindow(SGMB_Context* c, int width, int height)
{
    if (!c) return fail(SGMB_E_ARG, "NULL context");
    if (!((width == 5 && height == 5) || (width == 9 && height == 7)))
        return fail(SGMB_E_UNSUPPORTED, "census window %dx%d is not supported (5x5 = reference, 9x7 = 64-bit extension)", width, height);
    if (c->configured && (width != c->censusW || height != c->censusH)) {
        // buffers and the aggregation layout depend on the descriptor size: configure again before the next match
        if (int rc = ensure_device(c)) return rc;
        CU(cudaDeviceSynchronize());
        free_config(c);
    }
    c->censusW = width; c->censusH = height;
    return SGMB_OK;
}

extern "C" int SGMB_SetPipeline(SGMB_Context* c, unsigned flags)
{
    if (!c) return fail(SGMB_E_ARG, "NULL context");
    if (flags != c->pipeline) for (auto& s : c->slots) drop_frame_graph(s);
    c->pipeline = flags;
    if (c->configured && (flags & SGMB_PIPE_TAPS)) {
        if (int rc = ensure_device(c)) return rc;
        return alloc_taps(c);
    }
    return SGMB_OK;
}

// ------------------------------------------------------------------------------------------------ one frame
// Enqueue the whole pipeline on the slot's stream.  dL/dR: device images; result left in slot.dispFinal
// (or dOut when given).  Returns the number of kernels launched through *launches.
//   planar: dL / dR are planar B,G,R frames [3][N]; the census kernel converts them to grey on the fly (into s.img[]).
static int enqueue_frame(SGMB_Context* c, Slot& s, const uint8_t* dL, const uint8_t* dR, float* dOut, bool timeAgg, int* launches,
                         bool planar = false)
{
    const int W = c->W, H = c->H, D = c->D;
    const bool taps = (c->pipeline & SGMB_PIPE_TAPS) != 0;
    const bool doSpeckle = (c->pipeline & SGMB_PIPE_SPECKLE) && c->opt.is_remove_speckles;
    const bool doMedian = (c->pipeline & SGMB_PIPE_MEDIAN) != 0;
    int nk = 0;
    // SGMB_TimeKernels: an event before the first launch and one after every launch (never while capturing a graph)
    KernelTimer* tm = c->capturing ? nullptr : s.timer;
    if (tm) tm->n = 0;
    auto mark = [&](const char* name) -> int {
        if (!tm || tm->n >= kMaxTimedKernels) return SGMB_OK;
        tm->name[tm->n] = name;
        CU(cudaEventRecord(tm->ev[++tm->n], s.stream));
        return SGMB_OK;
    };

    if (c->nEntries > 0) CU(cudaMemsetAsync(s.side, 0, (size_t)c->nEntries * c->Dp * sizeof(uint16_t), s.stream));
    if (doMedian) {
        // the wavefront's exchange rows are cleared every frame (a few hundred KB, long before they are used), so the tag
        // epoch can stay constant and the frame has no per-launch kernel argument: it can be replayed as a CUDA graph
        CU(cudaMemsetAsync(s.xchg, 0, median_xchg_bytes(W, H), s.stream));       // exchange rows + the row-group ticket
        s.medianEpoch = 0;
    }

    if (tm) CU(cudaEventRecord(tm->ev[0], s.stream));
    {   // K1 census
        CensusParams p{};
        p.img[0] = dL; p.img[1] = dR; p.left = s.censusL; p.right4 = s.censusR4;
        p.copyStride = c->copyStride; p.padF = c->padF; p.W = W; p.H = H;
        p.grey[0] = s.img[0]; p.grey[1] = s.img[1];
        p.pixL = s.pixL;
        p.wR = (c->greyFormula == SGMB_GREY_STB) ? 77u : 76u; p.wG = 150u; p.wB = 29u;
        dim3 grid((W + kCensusTileW - 1) / kCensusTileW, (H + kCensusTileH - 1) / kCensusTileH, 2);
        const int threads = kCensusThreads;
        if (c->descBytes == 4) {
            if (planar) sgm_census<5, 5, uint32_t, true><<<grid, threads, 0, s.stream>>>(p);
            else        sgm_census<5, 5, uint32_t, false><<<grid, threads, 0, s.stream>>>(p);
        } else {
            if (planar) sgm_census<9, 7, desc64_t, true><<<grid, threads, 0, s.stream>>>(p);
            else        sgm_census<9, 7, desc64_t, false><<<grid, threads, 0, s.stream>>>(p);
        }
        ++nk;
        if (int rc = mark("sgm_census")) return rc;
        if (planar) dL = s.img[0];       // K2 reads the left grey image (adaptive P2, SemiGlobalMatching.c:335)
    }
    {   // K2 aggregation
        AggParams p{};
        p.img = dL; p.censusL = s.censusL; p.pixL = s.pixL; p.censusR4 = s.censusR4; p.copyStride = (uint32_t)c->copyStride; p.padF = c->padF;
        p.planes = s.planes; p.planeStride = c->planeStride; p.side = reinterpret_cast<uint32_t*>(s.side);
        p.entryOf = c->entryOf; p.work = c->work; p.nIrregularWarps = c->nIrregularWarps; p.nRegularWarps = c->nRegularWarps;
        p.W = W; p.H = H; p.D = D; p.Dp = c->Dp; p.dmin = c->opt.min_disparity; p.p1x2 = c->p1x2;
        {   // largest in-image census cost (5x5: 25 bits, 9x7: 63 with the centre bit always equal) + largest penalty
            const int maxCost = c->descBytes == 8 ? 62 : 25;
            const int maxP2 = std::min(256, std::max((int)c->opt.p1, (int)c->opt.p2_init));
            p.wrapInterior = (maxCost + maxP2 > 255) ? 1 : 0;
        }
        memcpy(p.p2x2, c->p2x2, sizeof p.p2x2);
        const int warps = c->nIrregularWarps + c->nRegularWarps;
        const int blocks = (warps + kAggWarpsPerBlock - 1) / kAggWarpsPerBlock;
        const int threads = kAggWarpsPerBlock * 32;
        if (timeAgg) CU(cudaEventRecordWithFlags(s.evAgg0, s.stream, c->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
        // PAD = false when the disparity range fills every lane of every layout exactly (see aggregate.cuh)
        const bool pad = (D != 64 * c->NR) || (c->altLayout == 2);
#define SGM_AGG_LAUNCH(NRH, LPPH, NRV, LPPV, NRI, DT)                                                          \
    do {                                                                                                       \
        if (agg_paired_layout(NRH, NRV) != (c->wtaCPP >= 8))                                                   \
            return fail(SGMB_E_STATE, "aggregation layout and plane byte order of the WTA kernel disagree");  \
        if (pad) sgm_aggregate_paths<NRH, LPPH, NRV, LPPV, NRI, DT, true><<<blocks, threads, 0, s.stream>>>(p);  \
        else     sgm_aggregate_paths<NRH, LPPH, NRV, LPPV, NRI, DT, false><<<blocks, threads, 0, s.stream>>>(p); \
    } while (0)
        if (c->descBytes == 4) {
            if (c->NR == 1)                           SGM_AGG_LAUNCH(2, 16, 4, 8, 1, uint32_t);
            else if (c->NR == 2 && c->altLayout == 1) SGM_AGG_LAUNCH(4, 16, 4, 16, 2, uint32_t);
            else if (c->NR == 2)                      SGM_AGG_LAUNCH(4, 16, 8, 8, 2, uint32_t);
            else                                      SGM_AGG_LAUNCH(8, 16, 8, 16, 4, uint32_t);
        } else {
            if (c->NR == 1)      SGM_AGG_LAUNCH(2, 16, 4, 8, 1, desc64_t);
            else if (c->NR == 2) SGM_AGG_LAUNCH(4, 16, 4, 16, 2, desc64_t);
            else                 SGM_AGG_LAUNCH(8, 16, 8, 16, 4, desc64_t);
        }
#undef SGM_AGG_LAUNCH
        if (timeAgg) CU(cudaEventRecordWithFlags(s.evAgg1, s.stream, c->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
        ++nk;
        if (int rc = mark("sgm_aggregate_paths")) return rc;
    }
    float* lrOut = s.dispLR;
    {   // K3 plane sum + WTA + LR
        WtaParams p{};
        p.planes = s.planes; p.planeStride = c->planeStride; p.nPlanes = c->nDirs;
        p.side = s.side; p.entryOf = c->entryOf; p.hasSide = c->nEntries > 0;
        p.S = taps ? s.S : nullptr; p.dispLeftWta = taps ? s.dispLeftWta : nullptr; p.dispRight = taps ? s.dispRight : nullptr;
        p.dispOut = lrOut;
        p.W = W; p.H = H; p.D = D; p.Dp = c->Dp; p.dmin = c->opt.min_disparity;
        p.checkUnique = c->opt.is_check_unique; p.oneMinusRatio = 1 - c->opt.uniqueness_ratio;
        p.checkLR = c->opt.is_check_lr; p.lrThres = c->opt.lrcheck_thres;
        p.ringCols = c->wtaRingRows;
        p.rightRow = s.rightRow;
        p.records = s.wtaRecords;
        p.fullRows = c->wtaFullRows; p.segments = c->wtaSegments; p.rowDone = s.wtaRowDone;
        p.rowsPerSm = c->wtaRowsPerSm; p.pieces = c->wtaPieces; p.sched = s.wtaSched;
        const int wtaGrid = c->wtaFullRows + c->wtaPieces;
#define SGM_WTA_LAUNCH(CPP)                                                                                          \
    if (c->nDirs == 8) {                                                                                             \
        if (taps) sgm_reduce_wta_lr<CPP, 8, true><<<wtaGrid, WtaShape<CPP>::kThreads, c->wtaSmem, s.stream>>>(p);          \
        else      sgm_reduce_wta_lr<CPP, 8, false><<<wtaGrid, WtaShape<CPP>::kThreads, c->wtaSmem, s.stream>>>(p);         \
    } else {                                                                                                         \
        if (taps) sgm_reduce_wta_lr<CPP, 4, true><<<wtaGrid, WtaShape<CPP>::kThreads, c->wtaSmem, s.stream>>>(p);          \
        else      sgm_reduce_wta_lr<CPP, 4, false><<<wtaGrid, WtaShape<CPP>::kThreads, c->wtaSmem, s.stream>>>(p);         \
    }
        switch (c->wtaCPP) {
            case 1:  SGM_WTA_LAUNCH(1); break;
            case 2:  SGM_WTA_LAUNCH(2); break;
            case 4:  SGM_WTA_LAUNCH(4); break;
            case 8:  SGM_WTA_LAUNCH(8); break;
            default: SGM_WTA_LAUNCH(16); break;
        }
#undef SGM_WTA_LAUNCH
        ++nk;
        if (int rc = mark("sgm_reduce_wta_lr")) return rc;
    }
    float* cur = lrOut;
    const int32_t* lab = nullptr;
    if (doSpeckle) {
        nk += launch_speckle_labels(cur, s.labels, W, H, 1.0f, c->opt.min_speckle_area, s.stream, mark);
        lab = s.labels;
        if (!doMedian) {
            speckle_apply<<<((int)c->N + 255) / 256, 256, 0, s.stream>>>(cur, s.dispSpeckle, lab, lab + c->N, (int)c->N, c->opt.min_speckle_area);
            ++nk;
            if (int rc = mark("speckle_apply")) return rc;
            cur = s.dispSpeckle;
        }
    }
    if (doMedian) {
        // the component sizes are applied while the median's inputs are gathered; dispSpeckle is a tap
        // without taps nobody reads dispFinal: the wavefront writes the caller's buffer itself (it only ever writes `out`)
        float* fin = (dOut && !taps) ? dOut : s.dispFinal;
        nk += launch_median3_inplace(cur, lab, lab ? lab + c->N : nullptr, c->opt.min_speckle_area, (taps && lab) ? s.dispSpeckle : nullptr,
                                     s.medianPrep, fin, s.xchg, &s.medianEpoch, W, H, s.stream, mark);
        cur = fin;
    }
    if (dOut && cur != dOut) CU(cudaMemcpyAsync(dOut, cur, c->N * sizeof(float), cudaMemcpyDeviceToDevice, s.stream));
    CU(cudaGetLastError());
    if (launches) *launches = nk;
    return SGMB_OK;
}

// Device-side address of caller memory the last kernel of the frame may write directly: page-locked host memory
// (cudaHostAlloc / cudaHostRegister / SGMB_HostAlloc) when the frame ends with the median and no taps are kept.
static float* direct_output(SGMB_Context* c, float* out)
{
    if (!out || !(c->pipeline & SGMB_PIPE_MEDIAN) || (c->pipeline & SGMB_PIPE_TAPS)) return nullptr;
    static const bool off = getenv("SGM_B200_NO_DIRECT_OUT") != nullptr;
    if (off) return nullptr;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, out) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? static_cast<float*>(a.devicePointer) : nullptr;
}

// The frame on the slot's own image buffers: recorded into a graph on first use, replayed afterwards.  `direct` (from
// direct_output(), or NULL): the median wavefront writes there instead of dispFinal - the node's `out` argument of the
// executable graph is re-pointed when it differs from the previous call's - and *wroteDirect tells the caller to skip its copy.
static int launch_slot_frame(SGMB_Context* c, Slot& s, float* direct = nullptr, bool* wroteDirect = nullptr)
{
    static const bool noGraph = getenv("SGM_B200_NO_GRAPH") != nullptr;
    if (wroteDirect) *wroteDirect = false;
    if (noGraph) {
        if (wroteDirect) *wroteDirect = direct != nullptr;
        return enqueue_frame(c, s, s.img[0], s.img[1], direct, false, nullptr);
    }
    if (!s.frameExec) {
        CU(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
        c->capturing = true;
        int rc = enqueue_frame(c, s, s.img[0], s.img[1], nullptr, false, nullptr);
        c->capturing = false;
        const cudaError_t e = cudaStreamEndCapture(s.stream, &s.frameGraph);
        if (rc == SGMB_OK && e != cudaSuccess) rc = fail(SGMB_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
        if (rc == SGMB_OK) {
            const cudaError_t e2 = cudaGraphInstantiate(&s.frameExec, s.frameGraph, 0);
            if (e2 != cudaSuccess) { s.frameExec = nullptr; rc = fail(SGMB_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e2)); }
        }
        if (rc == SGMB_OK) {
            size_t n = 0;
            CU(cudaGraphGetNodes(s.frameGraph, nullptr, &n));
            std::vector<cudaGraphNode_t> nodes(n);
            CU(cudaGraphGetNodes(s.frameGraph, nodes.data(), &n));
            for (cudaGraphNode_t nd : nodes) {
                cudaGraphNodeType t;
                CU(cudaGraphNodeGetType(nd, &t));
                if (t != cudaGraphNodeTypeKernel) continue;
                cudaKernelNodeParams kp{};
                CU(cudaGraphKernelNodeGetParams(nd, &kp));
                if (kp.func == reinterpret_cast<void*>(&median_wavefront)) s.medianNode = nd;
            }
            s.frameOut = s.dispFinal;
        }
        if (rc) { drop_frame_graph(s); return rc; }
    }
    float* want = (direct && s.medianNode) ? direct : s.dispFinal;
    if (s.medianNode && want != s.frameOut) {
        cudaKernelNodeParams kp{};
        CU(cudaGraphKernelNodeGetParams(s.medianNode, &kp));
        void* args[kMedianWavefrontArgs];
        for (int i = 0; i < kMedianWavefrontArgs; ++i) args[i] = kp.kernelParams[i];
        args[kMedianWavefrontOutArg] = &want;
        kp.kernelParams = args;
        CU(cudaGraphExecKernelNodeSetParams(s.frameExec, s.medianNode, &kp));
        s.frameOut = want;
    }
    if (wroteDirect) *wroteDirect = want == direct && direct != nullptr;
    CU(cudaGraphLaunch(s.frameExec, s.stream));
    return SGMB_OK;
}

static float* frame_result(SGMB_Context* c, Slot& s)
{
    const bool doSpeckle = (c->pipeline & SGMB_PIPE_SPECKLE) && c->opt.is_remove_speckles;
    const bool doMedian = (c->pipeline & SGMB_PIPE_MEDIAN) != 0;
    return doMedian ? s.dispFinal : (doSpeckle ? s.dispSpeckle : s.dispLR);
}

// ------------------------------------------------------------------------------------------------ host <-> device copies
// Page-locked caller memory (cudaHostAlloc / cudaHostRegister / SGMB_HostAlloc) is read and written by the copy engine
// directly and cudaMemcpyAsync is truly asynchronous.  Pageable memory (malloc, static arrays: what the reference demo
// passes, main.c:25-26,81) makes cudaMemcpyAsync stage through the driver's own buffers and block the calling thread.
//   * A single frame (SGM_Match / SGMB_Match) hands pageable pointers to the driver as they are: measured at C2 that
//     is as fast as staging through buffers of our own (0.82 ms per call either way, profiles/r2_o_kernels.jsonl;
//     page-locked buffers: 0.68 ms).
//   * A batch would lose its pipelining to those blocking copies, so there pageable buffers are staged through
//     page-locked buffers owned by the slot: one memcpy + ONE asynchronous copy for both images, and the result is
//     copied out of the staging buffer when the slot is reused or the batch ends.
static bool host_is_pinned(const void* p)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}


// Result of the slot's previous frame still sitting in the staging buffer -> the caller's pageable buffer.
static int flush_pending_out(SGMB_Context* c, Slot& s)
{
    if (!s.pendingOut) return SGMB_OK;
    CU(cudaStreamSynchronize(s.stream));
    memcpy(s.pendingOut, s.stageOut, c->N * sizeof(float));
    s.pendingOut = nullptr;
    return SGMB_OK;
}

static int copy_in(SGMB_Context* c, Slot& s, const uint8_t* L, const uint8_t* R, bool stagePageable)
{
    if (int rc = flush_pending_out(c, s)) return rc;
    if (!stagePageable || (host_is_pinned(L) && host_is_pinned(R))) {
        CU(cudaMemcpyAsync(s.img[0], L, c->N, cudaMemcpyHostToDevice, s.stream));
        CU(cudaMemcpyAsync(s.img[1], R, c->N, cudaMemcpyHostToDevice, s.stream));
        return SGMB_OK;
    }
    if (!s.stageIn) CU(cudaHostAlloc(reinterpret_cast<void**>(&s.stageIn), 2 * c->imgStride, cudaHostAllocDefault));
    else CU(cudaEventSynchronize(s.evCopied));       // the previous copy out of stageIn must have finished
    memcpy(s.stageIn, L, c->N);
    memcpy(s.stageIn + c->imgStride, R, c->N);
    CU(cudaMemcpyAsync(s.img[0], s.stageIn, c->imgStride + c->N, cudaMemcpyHostToDevice, s.stream));
    CU(cudaEventRecord(s.evCopied, s.stream));
    return SGMB_OK;
}

// Enqueue the copy of the slot's result to `out`; for pageable `out` the last step (staging -> out) happens in
// flush_pending_out(), which every path calls before it returns to the caller or reuses the slot.
static int copy_out(SGMB_Context* c, Slot& s, float* out, bool stagePageable)
{
    if (!stagePageable || host_is_pinned(out)) {
        CU(cudaMemcpyAsync(out, frame_result(c, s), c->N * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        return SGMB_OK;
    }
    if (!s.stageOut) CU(cudaHostAlloc(reinterpret_cast<void**>(&s.stageOut), c->N * sizeof(float), cudaHostAllocDefault));
    CU(cudaMemcpyAsync(s.stageOut, frame_result(c, s), c->N * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    s.pendingOut = out;
    return SGMB_OK;
}

extern "C" int SGMB_Match(SGMB_Context* c, const uint8_t* L, const uint8_t* R, float* out)
{
    if (!c || !c->configured) return fail(SGMB_E_STATE, "SGMB_Match: context is not configured");   // SemiGlobalMatching.c:70-72
    if (!L || !R) return fail(SGMB_E_ARG, "SGMB_Match: NULL image");                                 // SemiGlobalMatching.c:73-75
    if (int rc = ensure_device(c)) return rc;
    Slot& s = c->slots[0];
    const int rc = [&]() -> int {
        CU(cudaEventRecord(s.evStart, s.stream));
        if (int rc = copy_in(c, s, L, R, false)) return rc;
        bool wrote = false;
        if (int rc = launch_slot_frame(c, s, direct_output(c, out), &wrote)) return rc;
        if (out && !wrote) if (int rc = copy_out(c, s, out, false)) return rc;
        CU(cudaEventRecord(s.evStop, s.stream));
        CU(cudaStreamSynchronize(s.stream));
        CU(cudaEventElapsedTime(&c->lastMs, s.evStart, s.evStop));
        return flush_pending_out(c, s);
    }();
    if (rc != SGMB_OK) { cudaStreamSynchronize(s.stream); s.pendingOut = nullptr; }
    return rc;
}

extern "C" int SGMB_MatchDevice(SGMB_Context* c, const uint8_t* dL, const uint8_t* dR, float* dOut, int sync)
{
    if (!c || !c->configured) return fail(SGMB_E_STATE, "SGMB_MatchDevice: context is not configured");
    if (!dL || !dR || !dOut) return fail(SGMB_E_ARG, "SGMB_MatchDevice: NULL pointer");
    if (int rc = ensure_device(c)) return rc;
    Slot& s = c->slots[0];
    CU(cudaEventRecord(s.evStart, s.stream));
    if (int rc = enqueue_frame(c, s, dL, dR, dOut, false, nullptr)) return rc;
    CU(cudaEventRecord(s.evStop, s.stream));
    if (sync) {
        CU(cudaStreamSynchronize(s.stream));
        CU(cudaEventElapsedTime(&c->lastMs, s.evStart, s.evStop));
    }
    return SGMB_OK;
}

extern "C" int SGMB_Synchronize(SGMB_Context* c)
{
    if (!c) return fail(SGMB_E_ARG, "NULL context");
    if (int rc = ensure_device(c)) return rc;
    for (auto& s : c->slots) CU(cudaStreamSynchronize(s.stream));
    return SGMB_OK;
}

// ------------------------------------------------------------------------------------------------ batches
static int run_batch(SGMB_Context* c, const uint8_t* const* Ls, const uint8_t* const* Rs, float* const* outs, int n, bool deviceMem)
{
    if (!c || !c->configured) return fail(SGMB_E_STATE, "batch: context is not configured");
    if (n < 0 || (n > 0 && (!Ls || !Rs || !outs))) return fail(SGMB_E_ARG, "batch: bad arguments");
    // every pointer is checked before anything is enqueued: a bad pair must not leave earlier pairs running
    for (int k = 0; k < n; ++k)
        if (!Ls[k] || !Rs[k] || !outs[k]) return fail(SGMB_E_ARG, "batch: NULL pointer at pair %d", k);
    if (int rc = ensure_device(c)) return rc;
    Slot& s0 = c->slots[0];
    // On a CUDA error in the middle of the batch the work already enqueued (kernels, asynchronous copies into the caller's
    // buffers) is drained before the caller gets control back.
    const int rc = [&]() -> int {
        CU(cudaEventRecord(s0.evStart, s0.stream));
        for (int j = 1; j < c->nslots && n > 0; ++j) CU(cudaStreamWaitEvent(c->slots[j].stream, s0.evStart, 0));
        for (int k = 0; k < n; ++k) {
            Slot& s = c->slots[k % c->nslots];
            if (deviceMem) {
                if (int rc = enqueue_frame(c, s, Ls[k], Rs[k], outs[k], false, nullptr)) return rc;
            } else {
                if (int rc = copy_in(c, s, Ls[k], Rs[k], true)) return rc;
                // a batch already hides the copy of one frame's result behind the kernels of the next; the direct write is kept
                // for contexts with a single slot (measured with the 8-GPU pool: 24.1 k pairs/s with copies, 22.5 k with direct writes)
                bool wrote = false;
                if (int rc = launch_slot_frame(c, s, c->nslots == 1 ? direct_output(c, outs[k]) : nullptr, &wrote)) return rc;
                if (!wrote) if (int rc = copy_out(c, s, outs[k], true)) return rc;
            }
        }
        for (int j = 1; j < c->nslots; ++j) {
            CU(cudaEventRecord(c->slots[j].evDone, c->slots[j].stream));
            CU(cudaStreamWaitEvent(s0.stream, c->slots[j].evDone, 0));
        }
        CU(cudaEventRecord(s0.evStop, s0.stream));
        CU(cudaStreamSynchronize(s0.stream));
        CU(cudaEventElapsedTime(&c->lastMs, s0.evStart, s0.evStop));
        for (auto& s : c->slots) if (int rc = flush_pending_out(c, s)) return rc;
        return SGMB_OK;
    }();
    if (rc != SGMB_OK) for (auto& s : c->slots) { cudaStreamSynchronize(s.stream); s.pendingOut = nullptr; }
    return rc;
}

extern "C" int SGMB_MatchBatch(SGMB_Context* c, const uint8_t* const* Ls, const uint8_t* const* Rs, float* const* outs, int n)
{
    return run_batch(c, Ls, Rs, outs, n, false);
}

extern "C" int SGMB_MatchBatchDevice(SGMB_Context* c, const uint8_t* const* Ls, const uint8_t* const* Rs, float* const* outs, int n)
{
    return run_batch(c, Ls, Rs, outs, n, true);
}

// The one statement of the batch sharding rule: device g of ndev takes the contiguous pairs [lo, hi) of n.
extern "C" int SGMB_ShardRange(int n, int ndev, int g, int* lo, int* hi)
{
    if (n < 0 || ndev < 1 || g < 0 || g >= ndev || !lo || !hi) return fail(SGMB_E_ARG, "SGMB_ShardRange: bad arguments");
    *lo = (int)((long long)n * g / ndev);
    *hi = (int)((long long)n * (g + 1) / ndev);
    return SGMB_OK;
}

extern "C" int SGMB_MatchBatchMultiGPU(const int* devices, int ndev, int slots, uint16_t width, uint16_t height,
                                       const SGMOption* option, unsigned flags, const uint8_t* const* Ls,
                                       const uint8_t* const* Rs, float* const* outs, int n)
{
    if (!devices || ndev < 1 || n < 0 || !option) return fail(SGMB_E_ARG, "SGMB_MatchBatchMultiGPU: bad arguments");
    std::vector<int> rcs(ndev, SGMB_OK);
    std::vector<std::string> msgs(ndev);
    std::vector<std::thread> th;
    for (int g = 0; g < ndev; ++g) {
        th.emplace_back([&, g]() {
            int lo = 0, hi = 0;
            SGMB_ShardRange(n, ndev, g, &lo, &hi);
            SGMB_Context* ctx = nullptr;
            int rc = SGMB_Create(&ctx, devices[g], slots);
            if (!rc) rc = SGMB_SetPipeline(ctx, flags);
            if (!rc) rc = SGMB_Configure(ctx, width, height, option);
            if (!rc && hi > lo) rc = SGMB_MatchBatch(ctx, Ls + lo, Rs + lo, outs + lo, hi - lo);
            if (rc) msgs[g] = g_err;
            rcs[g] = rc;
            SGMB_Destroy(ctx);
        });
    }
    for (auto& t : th) t.join();
    for (int g = 0; g < ndev; ++g)
        if (rcs[g]) return fail(rcs[g], "device %d: %s", devices[g], msgs[g].c_str());
    return SGMB_OK;
}

// ------------------------------------------------------------------------------------------------ persistent multi-GPU pool
// One context per device, kept alive between batches (SGMB_MatchBatchMultiGPU above builds and destroys its contexts on
// every call, which costs device allocations).  Same sharding rule: pair k of n goes to device k*ndev/n.
struct SGMB_Pool {
    std::vector<SGMB_Context*> ctx;
};

extern "C" int SGMB_PoolCreate(SGMB_Pool** out, const int* devices, int ndev, int slots_per_device)
{
    if (!out || !devices || ndev < 1) return fail(SGMB_E_ARG, "SGMB_PoolCreate: bad arguments");
    *out = nullptr;
    auto* p = new SGMB_Pool();
    for (int g = 0; g < ndev; ++g) {
        SGMB_Context* c = nullptr;
        const int rc = SGMB_Create(&c, devices[g], slots_per_device);
        if (rc) { for (auto* x : p->ctx) SGMB_Destroy(x); delete p; return rc; }
        p->ctx.push_back(c);
    }
    *out = p;
    return SGMB_OK;
}

extern "C" void SGMB_PoolDestroy(SGMB_Pool* p)
{
    if (!p) return;
    for (auto* c : p->ctx) SGMB_Destroy(c);
    delete p;
}

extern "C" int SGMB_PoolSize(SGMB_Pool* p) { return p ? (int)p->ctx.size() : 0; }

extern "C" SGMB_Context* SGMB_PoolContext(SGMB_Pool* p, int index)
{
    return (p && index >= 0 && index < (int)p->ctx.size()) ? p->ctx[index] : nullptr;
}

// Runs fn(context, device index) on one host thread per device and returns the first error.
template <typename F>
static int pool_for_each(SGMB_Pool* p, F fn)
{
    if (!p) return fail(SGMB_E_ARG, "NULL pool");
    const int ndev = (int)p->ctx.size();
    std::vector<int> rcs(ndev, SGMB_OK);
    std::vector<std::string> msgs(ndev);
    std::vector<std::thread> th;
    for (int g = 0; g < ndev; ++g)
        th.emplace_back([&, g]() {
            rcs[g] = fn(p->ctx[g], g);
            if (rcs[g]) msgs[g] = g_err;
        });
    for (auto& t : th) t.join();
    for (int g = 0; g < ndev; ++g)
        if (rcs[g]) return fail(rcs[g], "device %d: %s", p->ctx[g]->device, msgs[g].c_str());
    return SGMB_OK;
}

extern "C" int SGMB_PoolConfigure(SGMB_Pool* p, uint16_t width, uint16_t height, const SGMOption* option, unsigned pipeline_flags)
{
    return pool_for_each(p, [&](SGMB_Context* c, int) {
        int rc = SGMB_SetPipeline(c, pipeline_flags);
        if (!rc) rc = SGMB_Configure(c, width, height, option);
        return rc;
    });
}

extern "C" int SGMB_PoolMatchBatch(SGMB_Pool* p, const uint8_t* const* lefts, const uint8_t* const* rights, float* const* disps, int n)
{
    if (n < 0 || (n > 0 && (!lefts || !rights || !disps))) return fail(SGMB_E_ARG, "SGMB_PoolMatchBatch: bad arguments");
    if (!p) return fail(SGMB_E_ARG, "NULL pool");
    const int ndev = (int)p->ctx.size();
    return pool_for_each(p, [&](SGMB_Context* c, int g) {
        int lo = 0, hi = 0;
        SGMB_ShardRange(n, ndev, g, &lo, &hi);
        return hi > lo ? SGMB_MatchBatch(c, lefts + lo, rights + lo, disps + lo, hi - lo) : SGMB_OK;
    });
}

// ------------------------------------------------------------------------------------------------ frame formats (N3) and evaluation (N4)
extern "C" int SGMB_SetGreyFormula(SGMB_Context* c, int formula)
{
    if (!c) return fail(SGMB_E_ARG, "NULL context");
    if (formula != SGMB_GREY_BOARD && formula != SGMB_GREY_STB) return fail(SGMB_E_ARG, "SGMB_SetGreyFormula: unknown formula %d", formula);
    c->greyFormula = formula;
    return SGMB_OK;
}

static int launch_depth(SGMB_Context* c, const float* dDisp, float* dDepth, size_t n, float baseline, float fx, float doffs, cudaStream_t st)
{
    const float bf = baseline * fx;                       // float32 product, as numpy forms it (depth_image.py:163)
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    sgm_disparity_to_depth<<<std::max(blocks, 1), 256, 0, st>>>(dDisp, dDepth, n, bf, doffs);
    CU(cudaGetLastError());
    (void)c;
    return SGMB_OK;
}

// One frame in the board's layout: six planes left B,G,R, right B,G,R (zb/frame_buffer.h:29-41).  calib20 == NULL:
// the result is the disparity map; otherwise it is the 20-float wire calibration (cam0, cam1 row-major, doffs,
// baseline; HostScript_Server/stereo_calibration.py:177-194) and the result is depth (depth_image.py:138-165).
static int match_frame(SGMB_Context* c, const uint8_t* planes6, const float* calib20, float* out, bool deviceMem, int sync)
{
    if (!c || !c->configured) return fail(SGMB_E_STATE, "SGMB_MatchFrame: context is not configured");
    if (!planes6 || !out) return fail(SGMB_E_ARG, "SGMB_MatchFrame: NULL pointer");
    if (int rc = ensure_device(c)) return rc;
    Slot& s = c->slots[0];
    const size_t N = c->N;
    CU(cudaEventRecord(s.evStart, s.stream));
    const uint8_t* dPlanes = planes6;
    if (!deviceMem) {
        if (!s.framePlanes) CU(cudaMalloc(&s.framePlanes, 6 * N + 16));
        CU(cudaMemcpyAsync(s.framePlanes, planes6, 6 * N, cudaMemcpyHostToDevice, s.stream));
        dPlanes = s.framePlanes;
    }
    if (int rc = enqueue_frame(c, s, dPlanes, dPlanes + 3 * N, nullptr, false, nullptr, true)) return rc;
    const float* result = frame_result(c, s);
    if (calib20) {
        if (!s.depth) CU(cudaMalloc(&s.depth, N * sizeof(float)));
        if (int rc = launch_depth(c, result, s.depth, N, calib20[19], calib20[0], calib20[18], s.stream)) return rc;
        result = s.depth;
    }
    CU(cudaMemcpyAsync(out, result, N * sizeof(float), deviceMem ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s.stream));
    CU(cudaEventRecord(s.evStop, s.stream));
    if (sync) {
        CU(cudaStreamSynchronize(s.stream));
        CU(cudaEventElapsedTime(&c->lastMs, s.evStart, s.evStop));
    }
    return SGMB_OK;
}

extern "C" int SGMB_MatchFrame(SGMB_Context* c, const uint8_t* planes6, const float* calib20, float* out)
{
    return match_frame(c, planes6, calib20, out, false, 1);
}

extern "C" int SGMB_MatchFrameDevice(SGMB_Context* c, const uint8_t* d_planes6, const float* calib20, float* d_out, int sync)
{
    return match_frame(c, d_planes6, calib20, d_out, true, sync);
}

extern "C" size_t SGMB_DepthReplyBytes(uint16_t width, uint16_t height) { return 9 + (size_t)width * height * sizeof(float); }

// Reply message of the board (zb/tcp_perf_client.c:106-131, parsed by HostScript_Server/server.py:148-177):
// type byte 3, frame id (32-bit LE), width, height (16-bit LE), then width*height float32 rows.
extern "C" int SGMB_PackDepthReply(uint32_t frame_id, uint16_t width, uint16_t height, const float* depth, uint8_t* dst, size_t capacity)
{
    const size_t need = SGMB_DepthReplyBytes(width, height);
    if (!depth || !dst) return fail(SGMB_E_ARG, "SGMB_PackDepthReply: NULL pointer");
    if (capacity < need) return fail(SGMB_E_STATE, "SGMB_PackDepthReply: needs %zu bytes, got %zu", need, capacity);
    dst[0] = 3;
    for (int i = 0; i < 4; ++i) dst[1 + i] = (uint8_t)(frame_id >> (8 * i));
    dst[5] = (uint8_t)width; dst[6] = (uint8_t)(width >> 8);
    dst[7] = (uint8_t)height; dst[8] = (uint8_t)(height >> 8);
    memcpy(dst + 9, depth, need - 9);
    return SGMB_OK;
}

// Frame header sent by the server: '<BiHH' = type (1: with 80-byte calibration, 2: images only), sequence
// number, width, height (HostScript_Server/server.py:114; parsed at zb/tcp_perf_client.c:154-201).
extern "C" int SGMB_ParseFrameHeader(const uint8_t* bytes9, int* type, int32_t* seq, uint16_t* width, uint16_t* height, size_t* payload_bytes)
{
    if (!bytes9) return fail(SGMB_E_ARG, "SGMB_ParseFrameHeader: NULL pointer");
    const int t = bytes9[0];
    if (t != 1 && t != 2) return fail(SGMB_E_ARG, "SGMB_ParseFrameHeader: message type %d carries no frame", t);
    const uint32_t q = (uint32_t)bytes9[1] | ((uint32_t)bytes9[2] << 8) | ((uint32_t)bytes9[3] << 16) | ((uint32_t)bytes9[4] << 24);
    const uint16_t w = (uint16_t)(bytes9[5] | (bytes9[6] << 8)), h = (uint16_t)(bytes9[7] | (bytes9[8] << 8));
    if (type) *type = t;
    if (seq) *seq = (int32_t)q;
    if (width) *width = w;
    if (height) *height = h;
    if (payload_bytes) *payload_bytes = (t == 1 ? 80 : 0) + (size_t)6 * w * h;
    return SGMB_OK;
}

static int depth_convert(SGMB_Context* c, const float* disp, size_t n, float baseline, float fx, float doffs, float* depth, bool deviceMem)
{
    if (!c) return fail(SGMB_E_ARG, "NULL context");
    if ((!disp || !depth) && n) return fail(SGMB_E_ARG, "SGMB_DisparityToDepth: NULL pointer");
    if (n == 0) return SGMB_OK;
    if (int rc = ensure_device(c)) return rc;
    cudaStream_t st = c->slots[0].stream;
    if (deviceMem) {
        if (int rc = launch_depth(c, disp, depth, n, baseline, fx, doffs, st)) return rc;
        CU(cudaStreamSynchronize(st));
        return SGMB_OK;
    }
    float* d = nullptr;
    CU(cudaMalloc(&d, n * sizeof(float)));
    int rc = SGMB_OK;
    cudaError_t e = cudaMemcpyAsync(d, disp, n * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = launch_depth(c, d, d, n, baseline, fx, doffs, st);
    if (e == cudaSuccess && rc == SGMB_OK) e = cudaMemcpyAsync(depth, d, n * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return fail(SGMB_E_CUDA, "SGMB_DisparityToDepth: %s", cudaGetErrorString(e));
    return rc;
}

extern "C" int SGMB_DisparityToDepth(SGMB_Context* c, const float* disp, size_t n, float baseline, float fx, float doffs, float* depth)
{
    return depth_convert(c, disp, n, baseline, fx, doffs, depth, false);
}

extern "C" int SGMB_DisparityToDepthDevice(SGMB_Context* c, const float* d_disp, size_t n, float baseline, float fx, float doffs, float* d_depth)
{
    return depth_convert(c, d_disp, n, baseline, fx, doffs, d_depth, true);
}

static int compare_depth(SGMB_Context* c, const float* gt, const float* test, size_t n, float thresh, double* rmse, double* bpr,
                         long long* nValid, bool deviceMem)
{
    if (!c) return fail(SGMB_E_ARG, "NULL context");
    if ((!gt || !test) && n) return fail(SGMB_E_ARG, "SGMB_CompareDepth: NULL pointer");
    if (int rc = ensure_device(c)) return rc;
    cudaStream_t st = c->slots[0].stream;
    if (!c->cmpScratch) CU(cudaMalloc(&c->cmpScratch, (kCompareBlocks + 1) * sizeof(CompareAcc)));
    float* tmp = nullptr;
    const float *dG = gt, *dT = test;
    if (!deviceMem && n) {
        CU(cudaMalloc(&tmp, 2 * n * sizeof(float)));
        cudaError_t e = cudaMemcpyAsync(tmp, gt, n * sizeof(float), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tmp + n, test, n * sizeof(float), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { cudaFree(tmp); return fail(SGMB_E_CUDA, "SGMB_CompareDepth: %s", cudaGetErrorString(e)); }
        dG = tmp; dT = tmp + n;
    }
    sgm_compare_depth_partial<<<kCompareBlocks, kCompareThreads, 0, st>>>(dG, dT, n, thresh, c->cmpScratch);
    sgm_compare_depth_final<<<1, 32, 0, st>>>(c->cmpScratch, kCompareBlocks, c->cmpScratch + kCompareBlocks);
    CompareAcc acc{};
    cudaError_t e = cudaMemcpyAsync(&acc, c->cmpScratch + kCompareBlocks, sizeof acc, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(SGMB_E_CUDA, "SGMB_CompareDepth: %s", cudaGetErrorString(e));
    // depth_image.py:300-302: no valid pixel -> (nan, nan, 0)
    if (nValid) *nValid = (long long)acc.nValid;
    if (rmse) *rmse = acc.nValid ? sqrt(acc.sumSq / (double)acc.nValid) : NAN;
    if (bpr) *bpr = acc.nValid ? (double)acc.nBad / (double)acc.nValid : NAN;
    return SGMB_OK;
}

extern "C" int SGMB_CompareDepth(SGMB_Context* c, const float* gt, const float* test, size_t n, float abs_thresh, double* rmse,
                                 double* bpr, long long* n_valid)
{
    return compare_depth(c, gt, test, n, abs_thresh, rmse, bpr, n_valid, false);
}

extern "C" int SGMB_CompareDepthDevice(SGMB_Context* c, const float* d_gt, const float* d_test, size_t n, float abs_thresh, double* rmse,
                                       double* bpr, long long* n_valid)
{
    return compare_depth(c, d_gt, d_test, n, abs_thresh, rmse, bpr, n_valid, true);
}

// ------------------------------------------------------------------------------------------------ taps
extern "C" int SGMB_GetStage(SGMB_Context* c, int stage, void* dst, size_t bytes)
{
    if (!c || !c->configured) return fail(SGMB_E_STATE, "SGMB_GetStage: context is not configured");
    if (!dst) return fail(SGMB_E_ARG, "SGMB_GetStage: NULL destination");
    if (int rc = ensure_device(c)) return rc;
    Slot& s = c->slots[0];
    CU(cudaStreamSynchronize(s.stream));
    const size_t N = c->N;
    const bool taps = c->tapsAllocated && (c->pipeline & SGMB_PIPE_TAPS);
    const void* src = nullptr;
    size_t need = 0;
    switch (stage) {
        case SGMB_STAGE_CENSUS_LEFT:   src = s.censusL; need = N * c->descBytes; break;
        case SGMB_STAGE_CENSUS_RIGHT:  src = static_cast<const uint8_t*>(s.censusR4) + (size_t)c->padF * c->descBytes; need = N * c->descBytes; break;   // copy 0 is unshifted
        case SGMB_STAGE_AGGR:          src = taps ? s.S : nullptr; need = N * c->D * 2; break;
        case SGMB_STAGE_DISP_LEFT_WTA: src = taps ? s.dispLeftWta : nullptr; need = N * 4; break;
        case SGMB_STAGE_DISP_RIGHT:    src = taps ? s.dispRight : nullptr; need = N * 4; break;
        case SGMB_STAGE_DISP_LR:       src = s.dispLR; need = N * 4; break;
        case SGMB_STAGE_DISP_SPECKLE:
            if ((c->pipeline & SGMB_PIPE_SPECKLE) && c->opt.is_remove_speckles) src = (taps || !(c->pipeline & SGMB_PIPE_MEDIAN)) ? s.dispSpeckle : nullptr;
            else src = s.dispLR;
            need = N * 4; break;
        case SGMB_STAGE_DISP_FINAL:    src = frame_result(c, s); need = N * 4; break;
        case SGMB_STAGE_SPECKLE_LABELS: src = s.labels; need = 2 * N * 4; break;
        case SGMB_STAGE_GREY_LEFT:     src = s.img[0]; need = N; break;
        case SGMB_STAGE_GREY_RIGHT:    src = s.img[1]; need = N; break;
        default:
            if (stage >= SGMB_STAGE_PATH_PLANE_0 && stage < SGMB_STAGE_PATH_PLANE_0 + c->nDirs) {
                need = N * c->D;
                if (bytes != need) return fail(SGMB_E_STATE, "SGMB_GetStage: stage %d needs %zu bytes, got %zu", stage, need, bytes);
                const uint8_t* plane = s.planes + (size_t)(stage - SGMB_STAGE_PATH_PLANE_0) * c->planeStride;
                if (c->wtaCPP >= 8) {
                    // paired plane layout (aggregate.cuh): inside every unit of 8 disparities the bytes lie as 0,4,1,5,2,6,3,7
                    std::vector<uint8_t> tmp(N * (size_t)c->Dp);
                    CU(cudaMemcpy(tmp.data(), plane, tmp.size(), cudaMemcpyDeviceToHost));
                    uint8_t* out = static_cast<uint8_t*>(dst);
                    for (size_t px = 0; px < N; ++px)
                        for (int d = 0; d < c->D; ++d)
                            out[px * c->D + d] = tmp[px * c->Dp + (d & ~7) + 2 * (d & 3) + ((d >> 2) & 1)];
                    return SGMB_OK;
                }
                CU(cudaMemcpy2D(dst, c->D, plane, c->Dp, c->D, N, cudaMemcpyDeviceToHost));
                return SGMB_OK;
            }
            return fail(SGMB_E_ARG, "SGMB_GetStage: unknown stage %d", stage);
    }
    if (!src) return fail(SGMB_E_STATE, "SGMB_GetStage: stage %d is only retained with SGMB_PIPE_TAPS", stage);
    if (bytes != need) return fail(SGMB_E_STATE, "SGMB_GetStage: stage %d needs %zu bytes, got %zu", stage, need, bytes);
    CU(cudaMemcpy(dst, src, need, cudaMemcpyDeviceToHost));
    return SGMB_OK;
}

extern "C" int SGMB_HostAlloc(void** out, size_t bytes)
{
    if (!out) return fail(SGMB_E_ARG, "NULL");
    CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return SGMB_OK;
}

extern "C" void SGMB_HostFree(void* p) { if (p) cudaFreeHost(p); }

// Page-lock memory the caller already owns (malloc, static arrays - what main.c:25-26,81 passes), so that SGM_Match reads and
// writes it without the driver's staging copies.  The range must be unregistered before it is freed.
extern "C" int SGMB_HostRegister(void* p, size_t bytes)
{
    if (!p || !bytes) return fail(SGMB_E_ARG, "SGMB_HostRegister: NULL pointer or zero size");
    CU(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    return SGMB_OK;
}

extern "C" int SGMB_HostUnregister(void* p)
{
    if (!p) return fail(SGMB_E_ARG, "SGMB_HostUnregister: NULL pointer");
    CU(cudaHostUnregister(p));
    return SGMB_OK;
}

// ------------------------------------------------------------------------------------------------ introspection
extern "C" int SGMB_KernelLaunchesPerFrame(SGMB_Context* c)
{
    if (!c || !c->configured) return fail(SGMB_E_STATE, "not configured");
    const bool doSpeckle = (c->pipeline & SGMB_PIPE_SPECKLE) && c->opt.is_remove_speckles;
    const bool doMedian = (c->pipeline & SGMB_PIPE_MEDIAN) != 0;
    return 3 + (doSpeckle ? kSpeckleLabelLaunches + (doMedian ? 0 : 1) : 0) + (doMedian ? kMedianLaunches : 0);
}

extern "C" double SGMB_ModelBytesPerFrame(SGMB_Context* c)
{
    if (!c || !c->configured) return 0.0;
    return (double)c->N * (4.0 * c->nDirs * c->D + 6.0);
}

extern "C" double SGMB_PlanBytesPerFrame(SGMB_Context* c)
{
    if (!c || !c->configured) return 0.0;
    // planes written once by K2 and read once by K3 (Dp bytes per pixel and direction), census/images/disparity per pixel
    return (double)c->N * (2.0 * c->nDirs * c->Dp + 2 + 4 + 8 * 4 + 8 * 4 + 4);
}

extern "C" float SGMB_LastDeviceMs(SGMB_Context* c) { return c ? c->lastMs : -1.f; }

__global__ void sgm_flush_l2(uint4* buf, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] = make_uint4(i, 0, 0, 0);
}

extern "C" int SGMB_TimeDevice(SGMB_Context* c, const uint8_t* dL, const uint8_t* dR, float* dOut, int warmup, int iters,
                               int flush_l2, float* frame_ms, float* agg_ms)
{
    if (!c || !c->configured) return fail(SGMB_E_STATE, "SGMB_TimeDevice: context is not configured");
    if (!dL || !dR || !dOut || iters < 1 || warmup < 0) return fail(SGMB_E_ARG, "SGMB_TimeDevice: bad arguments");
    if (int rc = ensure_device(c)) return rc;
    Slot& s = c->slots[0];
    if (flush_l2 && !c->flushBuf) {
        c->flushBytes = (size_t)256 << 20;   // 2x the 126 MB L2
        CU(cudaMalloc(&c->flushBuf, c->flushBytes));
    }
    for (int it = -warmup; it < iters; ++it) {
        if (flush_l2) sgm_flush_l2<<<148 * 8, 256, 0, s.stream>>>(reinterpret_cast<uint4*>(c->flushBuf), c->flushBytes / 16);
        CU(cudaEventRecord(s.evStart, s.stream));
        if (int rc = enqueue_frame(c, s, dL, dR, dOut, true, nullptr)) return rc;
        CU(cudaEventRecord(s.evStop, s.stream));
        CU(cudaStreamSynchronize(s.stream));
        if (it >= 0) {
            if (frame_ms) CU(cudaEventElapsedTime(&frame_ms[it], s.evStart, s.evStop));
            if (agg_ms) CU(cudaEventElapsedTime(&agg_ms[it], s.evAgg0, s.evAgg1));
        }
    }
    return SGMB_OK;
}

// Per-kernel durations of one frame: events around every launch of the current pipeline (direct launches, no graph),
// averaged over `iters` frames after `warmup` untimed ones.  Returns the number of kernels; names via SGMB_KernelName.
extern "C" int SGMB_TimeKernels(SGMB_Context* c, const uint8_t* dL, const uint8_t* dR, float* dOut, int warmup, int iters,
                                float* kernel_ms, int capacity)
{
    if (!c || !c->configured) return fail(SGMB_E_STATE, "SGMB_TimeKernels: context is not configured");
    if (!dL || !dR || !dOut || iters < 1 || warmup < 0 || !kernel_ms || capacity < 1) return fail(SGMB_E_ARG, "SGMB_TimeKernels: bad arguments");
    if (int rc = ensure_device(c)) return rc;
    Slot& s = c->slots[0];
    KernelTimer tm;
    int rc = SGMB_OK, n = 0;
    for (auto& e : tm.ev) if (cudaEventCreate(&e) != cudaSuccess) rc = fail(SGMB_E_CUDA, "cudaEventCreate failed");
    std::vector<double> sum(kMaxTimedKernels, 0.0);
    s.timer = &tm;
    for (int it = -warmup; it < iters && rc == SGMB_OK; ++it) {
        rc = enqueue_frame(c, s, dL, dR, dOut, false, nullptr);
        if (rc == SGMB_OK && cudaStreamSynchronize(s.stream) != cudaSuccess) rc = fail(SGMB_E_CUDA, "SGMB_TimeKernels: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc != SGMB_OK || it < 0) continue;
        n = tm.n;
        for (int k = 0; k < n; ++k) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, tm.ev[k], tm.ev[k + 1]);
            sum[k] += ms;
        }
    }
    s.timer = nullptr;
    for (int k = 0; k < kMaxTimedKernels; ++k) c->timedName[k] = k < n ? tm.name[k] : nullptr;
    for (auto& e : tm.ev) if (e) cudaEventDestroy(e);
    if (rc != SGMB_OK) return rc;
    for (int k = 0; k < n && k < capacity; ++k) kernel_ms[k] = (float)(sum[k] / iters);
    return n;
}

extern "C" const char* SGMB_KernelName(SGMB_Context* c, int index)
{
    return (c && index >= 0 && index < kMaxTimedKernels && c->timedName[index]) ? c->timedName[index] : "";
}

// Enqueue `iters` frames back to back on slot 0 (device-resident input and output, no host synchronisation in
// between) and time the whole region with CUDA events on that stream; optionally also every aggregation launch.
// The frames are recorded into ONE CUDA graph (kernel nodes, the memsets and - when asked for - external event records
// around every aggregation kernel) and replayed with a single launch: the launch-bound inner loop of a device-resident
// batch.  `replays` > 1 launches the same graph that many times, each replay timed on its own (replay_ms[r]).
static int run_device(SGMB_Context* c, const uint8_t* dL, const uint8_t* dR, float* dOut, int iters, int replays,
                      float* replay_ms, float* agg_ms)
{
    if (!c || !c->configured) return fail(SGMB_E_STATE, "SGMB_RunDevice: context is not configured");
    if (!dL || !dR || !dOut || iters < 1 || replays < 1) return fail(SGMB_E_ARG, "SGMB_RunDevice: bad arguments");
    if (int rc = ensure_device(c)) return rc;
    Slot& s = c->slots[0];
    std::vector<cudaEvent_t> ev(agg_ms ? 2 * (size_t)iters : 0, nullptr);
    const cudaEvent_t keep0 = s.evAgg0, keep1 = s.evAgg1;
    const bool useGraph = !getenv("SGM_B200_NO_GRAPH");
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool capturing = false;
    auto enqueue_all = [&]() -> int {
        for (int it = 0; it < iters; ++it) {
            if (agg_ms) { s.evAgg0 = ev[2 * it]; s.evAgg1 = ev[2 * it + 1]; }
            if (int rc = enqueue_frame(c, s, dL, dR, dOut, agg_ms != nullptr, nullptr)) return rc;
        }
        return SGMB_OK;
    };
    const int rc = [&]() -> int {
        for (auto& e : ev) CU(cudaEventCreate(&e));
        if (useGraph) {
            CU(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
            capturing = c->capturing = true;
            const int erc = enqueue_all();
            c->capturing = false;
            const cudaError_t e = cudaStreamEndCapture(s.stream, &graph);
            capturing = false;
            if (erc) return erc;
            if (e != cudaSuccess) return fail(SGMB_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
            CU(cudaGraphInstantiate(&exec, graph, 0));
            CU(cudaGraphUpload(exec, s.stream));
            CU(cudaStreamSynchronize(s.stream));
        }
        for (int r = 0; r < replays; ++r) {
            CU(cudaEventRecord(s.evStart, s.stream));
            if (useGraph) CU(cudaGraphLaunch(exec, s.stream));
            else if (int erc = enqueue_all()) return erc;
            CU(cudaEventRecord(s.evStop, s.stream));
            CU(cudaStreamSynchronize(s.stream));
            CU(cudaEventElapsedTime(&c->lastMs, s.evStart, s.evStop));
            if (replay_ms) replay_ms[r] = c->lastMs;
        }
        if (agg_ms) for (int it = 0; it < iters; ++it) CU(cudaEventElapsedTime(&agg_ms[it], ev[2 * it], ev[2 * it + 1]));
        return SGMB_OK;
    }();
    // every exit path: leave capture mode, restore the slot's own events, release what was created
    if (capturing) { c->capturing = false; cudaGraph_t g = nullptr; cudaStreamEndCapture(s.stream, &g); if (g) cudaGraphDestroy(g); }
    s.evAgg0 = keep0; s.evAgg1 = keep1;
    if (rc != SGMB_OK) cudaStreamSynchronize(s.stream);
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    return rc;
}

extern "C" int SGMB_RunDevice(SGMB_Context* c, const uint8_t* dL, const uint8_t* dR, float* dOut, int iters,
                              float* total_ms, float* agg_ms /* [iters] or NULL */)
{
    float ms = 0.f;
    const int rc = run_device(c, dL, dR, dOut, iters, 1, &ms, agg_ms);
    if (rc == SGMB_OK && total_ms) *total_ms = ms;
    return rc;
}

extern "C" int SGMB_RunDeviceReplays(SGMB_Context* c, const uint8_t* dL, const uint8_t* dR, float* dOut, int iters, int replays,
                                     float* replay_ms /* [replays] */, float* agg_ms /* [iters] of the last replay, or NULL */)
{
    if (!replay_ms) return fail(SGMB_E_ARG, "SGMB_RunDeviceReplays: bad arguments");
    return run_device(c, dL, dR, dOut, iters, replays, replay_ms, agg_ms);
}

// ------------------------------------------------------------------------------------------------ reference API
static SGMB_Context* g_ctx = nullptr;
static int g_device = -1;
static int g_censusW = 0, g_censusH = 0;      // 0: not set -> env SGM_B200_CENSUS ("9x7") or 5x5
static std::mutex g_mu;

extern "C" SGMB_Context* SGMB_GlobalContext(void) { return g_ctx; }

extern "C" int SGMB_SetGlobalDevice(int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx && g_ctx->device != device) { SGMB_Destroy(g_ctx); g_ctx = nullptr; }
    g_device = device;
    return SGMB_OK;
}

extern "C" int SGMB_SetGlobalCensusWindow(int width, int height)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!((width == 5 && height == 5) || (width == 9 && height == 7)))
        return fail(SGMB_E_UNSUPPORTED, "census window %dx%d is not supported", width, height);
    g_censusW = width; g_censusH = height;       // applied by the next SGM_Initialize / SGM_Reset
    return SGMB_OK;
}

// replaces SemiGlobalMatching.c:37-66
extern "C" bool SGM_Initialize(uint16_t width, uint16_t height, const SGMOption* option)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!option) { fail(SGMB_E_ARG, "SGM_Initialize: NULL option"); return false; }
    if (!g_ctx) {
        int dev = g_device;
        if (dev < 0) { const char* e = getenv("SGM_B200_DEVICE"); dev = e ? atoi(e) : 0; }
        if (SGMB_Create(&g_ctx, dev, 1) != SGMB_OK) { g_ctx = nullptr; return false; }
    }
    {
        int cw = g_censusW, ch = g_censusH;
        if (cw == 0) {
            const char* e = getenv("SGM_B200_CENSUS");
            if (!e || sscanf(e, "%dx%d", &cw, &ch) != 2) { cw = 5; ch = 5; }
        }
        if (SGMB_SetCensusWindow(g_ctx, cw, ch) != SGMB_OK) return false;
    }
    return SGMB_Configure(g_ctx, width, height, option) == SGMB_OK;
}

// replaces SemiGlobalMatching.c:128-132
extern "C" bool SGM_Reset(uint16_t width, uint16_t height, const SGMOption* option)
{
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_ctx) g_ctx->configured = false;
    }
    return SGM_Initialize(width, height, option);
}

// replaces SemiGlobalMatching.c:68-125
extern "C" bool SGM_Match(const uint8_t* img_left, const uint8_t* img_right, float* disp_left)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) { fail(SGMB_E_STATE, "SGM_Match: SGM_Initialize has not succeeded"); return false; }
    return SGMB_Match(g_ctx, img_left, img_right, disp_left) == SGMB_OK;
}
