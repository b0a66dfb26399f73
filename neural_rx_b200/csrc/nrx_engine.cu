// nrx_engine.cu — host side of libnrx_b200.so: weight packing, workspace layout, launch sequence
// and the C ABI declared in include/nrx_b200.h.
//
// build: see neural_rx_b200/build.py
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC ...
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nrx_b200.h"
#include "nrx_kernels.cuh"
#include "nrx_stack.cuh"
#ifdef NRX_EXPERIMENTAL_PLANS
#include "nrx_stack_pair.cuh"
#include "nrx_stack_tm.cuh"
#endif
#include "nrx_stack_ws.cuh"
#include "nrx_agg_ws.cuh"

using namespace nrx;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define NRX_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t err_ = (expr);                                                          \
        if (err_ != cudaSuccess)                                                            \
            return fail(NRX_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(err_));    \
    } while (0)

struct SepLayer {          // one SeparableConv2D position of a stack (n_stacks copies for Var-IO)
    uint8_t* blob = nullptr;
    uint32_t blob_bytes = 0;
    int kpad = 0, npad = 0, n_stacks = 1;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// memcpy between pageable and pinned host memory on a few threads (one core moves ~10 GB/s, the
// staged path of nrx_forward_host moves 80 MB per 30-slot step)
void par_memcpy(void* dst, const void* src, size_t bytes) {
    const unsigned hw = std::thread::hardware_concurrency();
    size_t n = bytes / (size_t(2) << 20);                      // at least 2 MB per thread
    if (n > 4) n = 4;
    if (hw && n > hw) n = hw;
    if (n < 2) { memcpy(dst, src, bytes); return; }
    const size_t part = align_up((bytes + n - 1) / n, 4096);   // => at most n parts
    std::thread th[3];
    int k = 0;
    for (size_t off = part; off < bytes; off += part) {
        const size_t len = bytes - off < part ? bytes - off : part;
        th[k++] = std::thread([=] { memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len); });
    }
    memcpy(dst, src, part);
    for (int i = 0; i < k; ++i) th[i].join();
}

// fp16 K-major SWIZZLE_128B image of W^T: rows n (NPAD), slabs of 64 k.  kmap[c] = K index of input c.
void pack_pw(uint8_t* img, const float* w, int cin, int cout, int npad, const std::vector<int>& kmap, int n_off = 0) {
    for (int c = 0; c < cin; ++c)
        for (int n = 0; n < cout; ++n) {
            const int k = kmap[c], row = n + n_off;
            *reinterpret_cast<__half*>(img + size_t(k / 64) * npad * 128 + sw128_offset(row, k % 64)) =
                __float2half(w[size_t(c) * cout + n]);
        }
}

}  // namespace

struct nrx_engine {
    nrx_model_desc d{};
    int device = 0, num_sms = 148;
    int num_it = 1, slots_per_pass = 0;
    int cin0 = 0;                                   // 4N + 2
    std::vector<SepLayer> init_layers;              // 3
    std::vector<std::vector<SepLayer>> upd_layers;  // [it][3]
    std::vector<uint8_t*> agg_blobs;                // [it]
    struct AggBias { float v[128]; };               // host copies of the message-MLP biases [b1 | b2]: kernel parameters of
    std::vector<AggBias> agg_bias;                  //   the pipelined two-user kernel (nrx_agg_ws.cuh)
    int agg_pipelined = 1;                          // 0: nrx_agg_kernel<2> also for two users (cross-check, NRX_OPT_AGG_PIPELINED)
    int stack_balanced = 1;                         // 0: equal chunks per plane (choose_chunks) instead of balanced CTA ranges
    uint8_t* readout_blob = nullptr;                // [n_io] heads
    uint8_t* stack_init_blob = nullptr;             // fused StateInit stacks [n_io]
    std::vector<uint8_t*> stack_upd_blobs;          // fused UpdateState stack per iteration
    struct StackBias { float v[320]; };             // host copies of the stack biases [128 | 128 | 64]: kernel parameters of
    std::vector<StackBias> init_bias, upd_bias;     //   the pipelined stack kernel (per StateInit stack / per iteration)
#ifdef NRX_EXPERIMENTAL_PLANS
    uint8_t* pair_init_blob = nullptr;              // CTA-pair kernels: [n_io][2 ranks] half-weight images
    std::vector<uint8_t*> pair_upd_blobs;           // [it] -> [2 ranks]
    struct TmConsts { uint32_t tap[3][64][9]; float bias[320]; };
    std::vector<TmConsts> tm_consts;                // [it] taps / biases of the TMEM-resident stack kernel (plan 4)
    std::vector<uint8_t*> tm_blobs;                 // [it] its pointwise B images (output channels in fragment order)
#endif
    int skip_inactive = 0;                          // nrx_set_skip_inactive: planes of inactive users are not computed
    int fused = 6;                                  // 6 (default): serial StateInit kernel + pipelined UpdateState kernels;
                                                    // 5: both pipelined; 4: TMEM-resident UpdateState stacks (nrx_stack_tm.cuh);
                                                    // 1: fused stacks + aggregation kernel, 2: fused stacks with the message
                                                    // MLP in their tail (two users only), 3: CTA-pair stack kernels
                                                    // (experimental), 0: layer-per-kernel
    int32_t* nn_index = nullptr;
    FoccEntry* focc = nullptr;
    float* pos_enc = nullptr;
    int32_t* data_index = nullptr;
    int n_pilot_slots = 0;
    // Aerial / TensorRT-shaped entry point (nrx_set_aerial_dmrs)
    int32_t* nn_prb = nullptr;                      // [U][F*T] ordinal of the nearest non-zero pilot (per-PRB rule)
    float* pos_enc_aerial = nullptr;                // [U][F][T][2]
    int aerial_pilots = 0;                          // non-zero pilots per user
    int64_t mac_fixed[NRX_MAX_IO] = {0};            // StateInit + readouts per head
    int64_t mac_per_it = 0;
    // per-kernel event timing (nrx_set_profiling)
    bool profiling = false;
    struct Span { int cls; cudaEvent_t a, b; };
    static constexpr size_t kMaxSpans = 1 << 16;    // launches recorded between two nrx_get_profile calls
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    // host-call staging (nrx_forward_host)
    static constexpr int kRing = 3;                 // chunks in flight in nrx_forward_host
    int host_chunk = 0;                             // slots per pipeline chunk (0 = default)
    cudaStream_t stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[kRing] = {}, ev_comp[kRing] = {}, ev_d2h[kRing] = {};
    // asynchronous host calls (nrx_forward_host_async): the chunk ring runs on across calls, every call has a ticket
    static constexpr int kCalls = 4;                // calls that may be in flight
    uint64_t chunk_seq = 0;                         // chunks enqueued so far (ring slot = chunk_seq % kRing)
    uint64_t call_seq = 0;                          // tickets handed out so far
    cudaEvent_t ev_call[kCalls] = {};               // recorded after the last device-to-host copy of a call
    size_t arena_chunk_slots = 0;                   // slots per chunk the ring buffers are sized for
    int arena_outs = 0;                             // output mask the ring buffers are sized for
    void* h_pin = nullptr;
    size_t h_pin_bytes = 0;
    void* d_io = nullptr;
    size_t d_io_bytes = 0;
    void* d_ws = nullptr;
    size_t d_ws_bytes = 0;
};

namespace {

int host_streams(nrx_engine* e);

struct Workspace {
    size_t partial, plist, z0, h1, h2, abuf, abuf2, sbuf, sbuf2, total;
};

int pass_slots(const nrx_engine* e, int batch) {
    int s = e->slots_per_pass <= 0 ? batch : e->slots_per_pass;
    return s > batch ? batch : s;
}

Workspace layout(const nrx_engine* e, int batch) {
    const size_t P = size_t(pass_slots(e, batch)) * e->d.max_num_tx * e->d.num_subcarriers * kT;
    Workspace w{};
    size_t off = 0;
    w.partial = off; off = align_up(off + size_t(batch) * kPowerParts * 4, 256);
    w.plist = off;   off = align_up(off + (size_t(pass_slots(e, batch)) * e->d.max_num_tx + 1) * 4, 256);   // active planes of a pass
    w.z0 = off;      off = align_up(off + P * 32 * 2, 256);
    // hidden activations exist in HBM only in the layer-per-kernel plan; the fused plans keep them on chip
    const size_t hid = e->fused == 0 ? P * 128 * 2 : 0;
    w.h1 = off;      off = align_up(off + hid, 256);
    w.h2 = off;      off = align_up(off + hid, 256);
    w.abuf = off;    off = align_up(off + P * 64 * 2, 256);
    w.sbuf = off;    off = align_up(off + P * 64 * 2, 256);
    w.sbuf2 = off;   off = align_up(off + P * 64 * 2, 256);
    w.abuf2 = off;   off = align_up(off + P * 64 * 2, 256);
    w.total = off;
    return w;
}

template <typename K>
cudaError_t set_smem(K kernel, int bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

std::vector<int> identity_map(int n) {
    std::vector<int> m(n);
    for (int i = 0; i < n; ++i) m[i] = i;
    return m;
}

// Build the device blob of one sep-conv layer position for all stacks.
int build_sep_layer(SepLayer& L, int n_stacks, const float* const* arrays, const int64_t* sizes, int first,
                    int stride, int cin, int cout, int kpad, int npad, const std::vector<int>& kmap) {
    const int KS = (kpad + 63) / 64;
    const uint32_t wpw = KS * npad * 128, wdw = 9 * kpad * 2, wb = npad * 4;
    L.kpad = kpad;
    L.npad = npad;
    L.n_stacks = n_stacks;
    L.blob_bytes = wpw + wdw + wb;
    std::vector<uint8_t> host(size_t(L.blob_bytes) * n_stacks, 0);
    for (int s = 0; s < n_stacks; ++s) {
        const int i = first + s * stride;
        if (sizes[i] != int64_t(9) * cin || sizes[i + 1] != int64_t(cin) * cout || sizes[i + 2] != cout)
            return fail(NRX_ERR_INVALID, "SeparableConv2D(%d->%d) expected at weight index %d", cin, cout, i);
        uint8_t* b = host.data() + size_t(s) * L.blob_bytes;
        pack_pw(b, arrays[i + 1], cin, cout, npad, kmap);
        __half* dw = reinterpret_cast<__half*>(b + wpw);
        for (int tap = 0; tap < 9; ++tap)                      // [3][3][cin][1] -> [9][KPAD]
            for (int c = 0; c < cin; ++c) dw[tap * kpad + kmap[c]] = __float2half(arrays[i][size_t(tap) * cin + c]);
        float* bias = reinterpret_cast<float*>(b + wpw + wdw);
        for (int n = 0; n < cout; ++n) bias[n] = arrays[i + 2][n];
    }
    NRX_CUDA(cudaMalloc(&L.blob, host.size()));
    NRX_CUDA(cudaMemcpy(L.blob, host.data(), host.size(), cudaMemcpyHostToDevice));
    return NRX_OK;
}

// Weight image of one fused stack (StackSmem<MODE> layout): three pointwise B images, three tap
// tables [9][K] fp16, biases fp32 [128 | 128 | 64].  `first` indexes the stack's first array
// (depthwise, pointwise, bias per layer).
// Message-MLP section of a stack blob: AggregateUserStates Dense(d_s->units_agg), Dense(units_agg->d_s)
// of the iteration that FOLLOWS the stack (arrays[first..first+3] = k1, b1, k2, b2).
template <int MODE>
void pack_stack_agg(uint8_t* b, const float* const* arrays, int first, int d_s, int units_agg) {
    using L = StackSmem<MODE>;
    pack_pw(b + L::oAggW1, arrays[first], d_s, units_agg, 64, identity_map(d_s));
    pack_pw(b + L::oAggW2, arrays[first + 2], units_agg, d_s, 64, identity_map(units_agg));
    float* bias = reinterpret_cast<float*>(b + L::oAggB);
    for (int n = 0; n < units_agg; ++n) bias[n] = arrays[first + 1][n];
    for (int n = 0; n < d_s; ++n) bias[64 + n] = arrays[first + 3][n];
}

#ifdef NRX_EXPERIMENTAL_PLANS
// Per-rank weight image of the CTA-pair stack kernel (StackPairSmem<MODE>): rank r holds output channels
// [r*N/2, (r+1)*N/2) of every pointwise matrix (rows of the B image), all taps and all biases.
template <int MODE>
void pack_stack_pair_blob(uint8_t* b, int rank, const float* const* arrays, int first, const int (&widths)[4],
                          const std::vector<int>& kmap1) {
    using L = StackPairSmem<MODE>;
    const int w_off[3] = {L::oPw1, L::oPw2, L::oPw3}, t_off[3] = {L::oTap1, L::oTap2, L::oTap3};
    const int kpad[3] = {L::KP1, 128, 128}, nh[3] = {64, 64, 32}, b_off[3] = {0, 128, 256};
    for (int l = 0; l < 3; ++l) {
        const int i = first + 3 * l, cin = widths[l], cout = widths[l + 1];
        const std::vector<int> km = l == 0 ? kmap1 : identity_map(cin);
        for (int c = 0; c < cin; ++c)
            for (int n = rank * nh[l]; n < (rank + 1) * nh[l] && n < cout; ++n) {
                const int k = km[c], row = n - rank * nh[l];
                *reinterpret_cast<__half*>(b + w_off[l] + size_t(k / 64) * nh[l] * 128 + sw128_offset(row, k % 64)) =
                    __float2half(arrays[i + 1][size_t(c) * cout + n]);
            }
        __half* dw = reinterpret_cast<__half*>(b + t_off[l]);
        for (int tap = 0; tap < 9; ++tap)
            for (int c = 0; c < cin; ++c) dw[tap * kpad[l] + km[c]] = __float2half(arrays[i][size_t(tap) * cin + c]);
        float* bias = reinterpret_cast<float*>(b + L::oBias) + b_off[l];
        for (int n = 0; n < cout; ++n) bias[n] = arrays[i + 2][n];
    }
}

#endif

template <int MODE>
int pack_stack_blob(uint8_t* b, const float* const* arrays, const int64_t* sizes, int first, const int (&widths)[4],
                    const std::vector<int>& kmap1) {
    using L = StackSmem<MODE>;
    const int w_off[3] = {L::oPw1, L::oPw2, L::oPw3}, t_off[3] = {L::oTap1, L::oTap2, L::oTap3};
    const int kpad[3] = {L::KP1, 128, 128}, npad[3] = {128, 128, 64}, b_off[3] = {0, 128, 256};
    for (int l = 0; l < 3; ++l) {
        const int i = first + 3 * l, cin = widths[l], cout = widths[l + 1];
        if (sizes[i] != int64_t(9) * cin || sizes[i + 1] != int64_t(cin) * cout || sizes[i + 2] != cout)
            return fail(NRX_ERR_INVALID, "SeparableConv2D(%d->%d) expected at weight index %d", cin, cout, i);
        const std::vector<int> km = l == 0 ? kmap1 : identity_map(cin);
        pack_pw(b + w_off[l], arrays[i + 1], cin, cout, npad[l], km);
        __half* dw = reinterpret_cast<__half*>(b + t_off[l]);
        for (int tap = 0; tap < 9; ++tap)
            for (int c = 0; c < cin; ++c) dw[tap * kpad[l] + km[c]] = __float2half(arrays[i][size_t(tap) * cin + c]);
        float* bias = reinterpret_cast<float*>(b + L::oBias) + b_off[l];
        for (int n = 0; n < cout; ++n) bias[n] = arrays[i + 2][n];
    }
    return NRX_OK;
}

cudaEvent_t take_event(nrx_engine* e) {
    if (!e->event_pool.empty()) {
        cudaEvent_t ev = e->event_pool.back();
        e->event_pool.pop_back();
        return ev;
    }
    cudaEvent_t ev = nullptr;
    cudaEventCreate(&ev);
    return ev;
}

// RAII bracket: records an event pair around one launch when profiling is on
struct Timed {
    nrx_engine* e;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    int cls;
    Timed(nrx_engine* e_, cudaStream_t st_, int cls_) : e(e_), st(st_), cls(cls_) {
        // no timing events inside a stream capture (they would become graph nodes), and a bounded record: a caller
        // that never collects the profile must not grow it without limit
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (e->profiling && e->spans.size() < nrx_engine::kMaxSpans &&
            cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) {
            a = take_event(e);
            b = take_event(e);
            cudaEventRecord(a, st);
        }
    }
    ~Timed() {
        if (a) {
            cudaEventRecord(b, st);
            e->spans.push_back({cls, a, b});
        }
    }
};

template <int KPAD, int NPAD, int MODE>
void launch_sep(nrx_engine* e, cudaStream_t st, SepParams p) {
    constexpr int cls = KPAD == 32 ? NRX_K_SEP_IN : MODE == kHidden ? NRX_K_SEP_HID
                        : MODE == kInitOut ? NRX_K_SEP_INIT_OUT : NRX_K_SEP_UPD_OUT;
    Timed t(e, st, cls);
    const int grid = p.num_tiles < 2 * e->num_sms ? p.num_tiles : 2 * e->num_sms;
    nrx_sepconv_kernel<KPAD, NPAD, MODE><<<grid, kThreads, SepSmem<KPAD, NPAD>::kTotal, st>>>(p);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
using TensorMapEncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr{};
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
            qr != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<TensorMapEncodeFn>(p);
    }();
    return fn;
}

// A tensor map is a pure function of (address, shape, box, swizzle): the forwards of an engine encode the same dozen maps
// over and over (workspace buffers, ping-pong states), so the encoded descriptors are kept in a small per-thread cache
// instead of calling the driver for every launch.
struct MapKey {
    const void* base;
    int planes, rows, ch, box_rows, swizzle;
    bool operator==(const MapKey& o) const {
        return base == o.base && planes == o.planes && rows == o.rows && ch == o.ch && box_rows == o.box_rows && swizzle == o.swizzle;
    }
};
bool cached_map(const MapKey& k, CUtensorMap* m, bool store) {
    thread_local std::vector<std::pair<MapKey, CUtensorMap>> cache;
    if (store) {
        if (cache.size() >= 64) cache.clear();
        cache.emplace_back(k, *m);
        return true;
    }
    for (const auto& kv : cache)
        if (kv.first == k) { *m = kv.second; return true; }
    return false;
}

// [planes][rows][64] fp16 activation tensor, box = box_rows rows x 128 B, 128-byte swizzle
int make_rows_map(CUtensorMap* m, const __half* base, int planes, int rows, int box_rows) {
    const MapKey key{base, planes, rows, 64, box_rows, 1};
    if (cached_map(key, m, false)) return NRX_OK;
    if (!tensor_map_encoder()) return fail(NRX_ERR_CUDA, "cuTensorMapEncodeTiled: driver entry point not found");
    const cuuint64_t dims[3] = {64, cuuint64_t(rows), cuuint64_t(planes)};
    const cuuint64_t strides[2] = {128, cuuint64_t(rows) * 128};
    const cuuint32_t box[3] = {64, cuuint32_t(box_rows), 1}, estr[3] = {1, 1, 1};
    const CUresult r = tensor_map_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NRX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(r));
    cached_map(key, m, true);
    return NRX_OK;
}
// [planes][rows][ch] fp16 activation tensor, box = box_rows rows x ch channels, no swizzle: the layer-1 window of
// the pipelined stack kernel (rows before / after the plane read zeros = the 'same' padding of the first layer)
int make_window_map(CUtensorMap* m, const __half* base, int planes, int rows, int ch, int box_rows) {
    const MapKey key{base, planes, rows, ch, box_rows, 0};
    if (cached_map(key, m, false)) return NRX_OK;
    if (!tensor_map_encoder()) return fail(NRX_ERR_CUDA, "cuTensorMapEncodeTiled: driver entry point not found");
    const cuuint64_t dims[3] = {cuuint64_t(ch), cuuint64_t(rows), cuuint64_t(planes)};
    const cuuint64_t strides[2] = {cuuint64_t(ch) * 2, cuuint64_t(rows) * ch * 2};
    const cuuint32_t box[3] = {cuuint32_t(ch), cuuint32_t(box_rows), 1}, estr[3] = {1, 1, 1};
    const CUresult r = tensor_map_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NRX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(r));
    cached_map(key, m, true);
    return NRX_OK;
}
// box = one subcarrier (14 rows) of a (slot, user) plane
int make_plane_map(CUtensorMap* m, const __half* base, int planes, int F) { return make_rows_map(m, base, planes, F * kT, kT); }

int launch_agg(nrx_engine* e, cudaStream_t st, const __half* s, __half* a, int it, const float* active,
               int U, int per_slot, int bp, int skip_idle = 0) {
    const uint8_t* blob = e->agg_blobs[it];
    AggParams ap{};
    ap.skip_idle = skip_idle;
    const int rc = make_rows_map(&ap.map_s, s, bp * U, per_slot, 128);
    if (rc) return rc;
    ap.sbuf = s; ap.abuf = a; ap.wblob = blob; ap.active_tx = active;
    ap.U = U; ap.rows_per_bu = per_slot;
    ap.tiles_per_b = (per_slot + 127) / 128;
    ap.num_tiles = ap.tiles_per_b * bp;
    Timed t(e, st, NRX_K_AGG);
    if (U == 2 && e->agg_pipelined) {
        AggWsParams wp{};
        wp.a = ap;
        if (const int rc2 = make_rows_map(&wp.map_a, a, bp * U, per_slot, 128)) return rc2;
        std::memcpy(wp.b1, e->agg_bias[it].v, sizeof(wp.b1));
        std::memcpy(wp.b2, e->agg_bias[it].v + 64, sizeof(wp.b2));
        nrx_agg_ws_kernel<<<ap.num_tiles < e->num_sms ? ap.num_tiles : e->num_sms, kAggWsThreads, kAggWsSmem, st>>>(wp);
        return NRX_OK;
    }
    const int cap = agg_ctas_per_sm(U) * e->num_sms;
    const int grid = ap.num_tiles < cap ? ap.num_tiles : cap;
    switch (U) {
        case 1: nrx_agg_kernel<1><<<grid, kAggThreads, agg_smem_bytes(1), st>>>(ap); break;
        case 2: nrx_agg_kernel<2><<<grid, kAggThreads, agg_smem_bytes(2), st>>>(ap); break;
        case 3: nrx_agg_kernel<3><<<grid, kAggThreads, agg_smem_bytes(3), st>>>(ap); break;
        default: nrx_agg_kernel<4><<<grid, kAggThreads, agg_smem_bytes(4), st>>>(ap); break;
    }
    return NRX_OK;
}

// Work distribution of an nrx_stack_kernel launch over `planes` planes: balanced CTA ranges (default) or equal chunks per
// plane.  Returns the grid size.
int set_stack_work(const nrx_engine* e, StackParams& kp, int planes, int F) {
    kp.num_planes = planes;
    if (e->stack_balanced) {
        kp.n_chunks = 0;
        kp.num_items = 0;
        const long long g = (long long)planes * F / kMinSeg;
        return int(g < 1 ? 1 : g > e->num_sms ? e->num_sms : g);
    }
    kp.n_chunks = choose_chunks(planes, F, e->num_sms);
    kp.num_items = kp.n_chunks * planes;
    return kp.num_items < e->num_sms ? kp.num_items : e->num_sms;
}

// plan 5: one launch of the warp-specialised, pipelined stack kernel (nrx_stack_ws.cuh)
template <int MODE>
int launch_stack_ws(nrx_engine* e, cudaStream_t st, StackParams kp, int planes, int F, const float* bias) {
    WsParams wp{};
    memcpy(wp.bias, bias, sizeof wp.bias);
    int rc;
    if (MODE == kStackInit) rc = make_window_map(&wp.map_a, kp.z0, planes, F * kT, 32, kWsWin * kT);
    else {
        rc = make_window_map(&wp.map_a, kp.a_in, planes, F * kT, 64, kWsWin * kT);
        if (!rc) rc = make_window_map(&wp.map_s, kp.s_in, planes, F * kT, 64, kWsWin * kT);
    }
    if (rc) return rc;
    int grid = set_stack_work(e, kp, planes, F);
    if (!e->stack_balanced) {
        kp.n_chunks = ws_choose_chunks(planes, F, e->num_sms);
        kp.num_items = kp.n_chunks * planes;
        grid = kp.num_items < e->num_sms ? kp.num_items : e->num_sms;
    }
    kp.sp_out = nullptr;
    kp.pair_agg = 0;
    wp.p = kp;
    Timed t(e, st, MODE == kStackInit ? NRX_K_STACK_INIT : NRX_K_STACK_UPD);
    nrx_stack_ws_kernel<MODE><<<grid, kWsThreads, WsSmem<MODE>::kTotal, st>>>(wp);
    return NRX_OK;
}

#ifdef NRX_EXPERIMENTAL_PLANS
// plan 4: one TMEM-resident UpdateState stack launch
int launch_stack_tm(nrx_engine* e, cudaStream_t st, int it, const __half* a_in, const __half* s_in, __half* s_out, int planes, int F) {
    TmParams tp{};
    int rc = make_plane_map(&tp.map_a, a_in, planes, F);
    if (!rc) rc = make_plane_map(&tp.map_s, s_in, planes, F);
    if (!rc) rc = make_plane_map(&tp.map_o, s_out, planes, F);
    if (rc) return rc;
    memcpy(tp.tap, e->tm_consts[it].tap, sizeof tp.tap);
    memcpy(tp.bias, e->tm_consts[it].bias, sizeof tp.bias);
    tp.wblob = e->tm_blobs[it];
    tp.F = F;
    tp.jobs_per_plane = tm_choose_jobs(planes, F, e->num_sms);
    tp.num_jobs = planes * tp.jobs_per_plane;
    tp.num_items = (tp.num_jobs + kTmSeqs - 1) / kTmSeqs;
    tp.steps_per_item = (F + tp.jobs_per_plane - 1) / tp.jobs_per_plane + kTmFill;
    const int grid = tp.num_items < e->num_sms ? tp.num_items : e->num_sms;
    Timed t(e, st, NRX_K_STACK_UPD);
    nrx_stack_tm_kernel<<<grid, kTmThreads, TmSmem::kTotal, st>>>(tp);
    return NRX_OK;
}

#endif

}  // namespace

extern "C" {

const char* nrx_last_error(void) { return g_last_error.c_str(); }
const char* nrx_version(void) { return "nrx_b200 0.1 (sm_100a)"; }

int nrx_destroy(nrx_engine* e) {
    if (!e) return NRX_OK;
    cudaSetDevice(e->device);
    for (auto& L : e->init_layers) cudaFree(L.blob);
    for (auto& it : e->upd_layers)
        for (auto& L : it) cudaFree(L.blob);
    for (auto* b : e->agg_blobs) cudaFree(b);
    cudaFree(e->readout_blob);
    cudaFree(e->stack_init_blob);
    for (auto* b : e->stack_upd_blobs) cudaFree(b);
#ifdef NRX_EXPERIMENTAL_PLANS
    for (auto* b : e->tm_blobs) cudaFree(b);
    cudaFree(e->pair_init_blob);
    for (auto* b : e->pair_upd_blobs) cudaFree(b);
#endif
    cudaFree(e->nn_index);
    cudaFree(e->focc);
    cudaFree(e->pos_enc);
    cudaFree(e->data_index);
    cudaFree(e->nn_prb);
    cudaFree(e->pos_enc_aerial);
    cudaFree(e->d_io);
    cudaFree(e->d_ws);
    if (e->h_pin) cudaFreeHost(e->h_pin);
    if (e->stream) {
        cudaStreamDestroy(e->stream);
        cudaStreamDestroy(e->s_h2d);
        cudaStreamDestroy(e->s_d2h);
        for (int r = 0; r < nrx_engine::kRing; ++r) {
            if (e->ev_h2d[r]) cudaEventDestroy(e->ev_h2d[r]);
            if (e->ev_comp[r]) cudaEventDestroy(e->ev_comp[r]);
            if (e->ev_d2h[r]) cudaEventDestroy(e->ev_d2h[r]);
        }
        for (int c = 0; c < nrx_engine::kCalls; ++c)
            if (e->ev_call[c]) cudaEventDestroy(e->ev_call[c]);
    }
    for (auto& s : e->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto ev : e->event_pool) cudaEventDestroy(ev);
    delete e;
    return NRX_OK;
}

int nrx_create(const nrx_model_desc* desc, const float* const* weight_arrays, const int64_t* weight_sizes,
               int32_t num_arrays, const float* pilots, const int32_t* nn_index, const float* pos_enc,
               const int32_t* data_index, int32_t device, nrx_engine** out) {
    if (!desc || !weight_arrays || !weight_sizes || !pilots || !nn_index || !pos_enc || !data_index || !out)
        return fail(NRX_ERR_INVALID, "nrx_create: null argument");
    const nrx_model_desc& d = *desc;
    if (d.num_ofdm_symbols != kT) return fail(NRX_ERR_UNSUPPORTED, "only 14-symbol slots are implemented");
    if (d.num_rx_ant < 1 || d.num_rx_ant > 7) return fail(NRX_ERR_UNSUPPORTED, "num_rx_ant must be in [1, 7]");
    if (d.max_num_tx < 1 || d.max_num_tx > kAggMaxU) return fail(NRX_ERR_UNSUPPORTED, "max_num_tx must be in [1, 4]");
    if (d.d_s < 4 || d.d_s > 60 || d.d_s % 4) return fail(NRX_ERR_UNSUPPORTED, "d_s must be a multiple of 4 in [4, 60]");
    for (int i = 0; i < 2; ++i)
        if (d.units_init[i] < 1 || d.units_init[i] > 128 || d.units_state[i] < 1 || d.units_state[i] > 128)
            return fail(NRX_ERR_UNSUPPORTED, "hidden sep-conv widths must be <= 128");
    if (d.units_agg < 1 || d.units_agg > 64) return fail(NRX_ERR_UNSUPPORTED, "units_agg must be <= 64");
    if (d.units_readout < 1 || d.units_readout > 128) return fail(NRX_ERR_UNSUPPORTED, "units_readout must be <= 128");
    if (d.n_io < 1 || d.n_io > NRX_MAX_IO) return fail(NRX_ERR_INVALID, "n_io out of range");
    for (int m = 0; m < d.n_io; ++m)
        if (d.io_bits[m] < 1 || d.io_bits[m] > 16) return fail(NRX_ERR_UNSUPPORTED, "LLR head width must be <= 16");
    if (d.num_it < 1) return fail(NRX_ERR_INVALID, "num_it must be >= 1");
    if (d.num_subcarriers > 65535) return fail(NRX_ERR_UNSUPPORTED, "num_subcarriers must be < 65536");
    if (d.num_subcarriers < 1 || d.focc_block < 1 || d.num_subcarriers % d.focc_block)
        return fail(NRX_ERR_INVALID, "num_subcarriers must be a positive multiple of focc_block");
    if (d.num_dmrs_symbols < 1 || d.num_dmrs_symbols > NRX_MAX_DMRS) return fail(NRX_ERR_INVALID, "bad DMRS symbol count");
    const int expected = d.n_io * 9 + d.num_it * (4 + 9) + d.n_io * 4 + 4;
    if (num_arrays != expected)
        return fail(NRX_ERR_INVALID, "weight list has %d arrays, architecture consumes %d", num_arrays, expected);

    NRX_CUDA(cudaSetDevice(device));
    nrx_engine* e = new nrx_engine();
    if (!tensor_map_encoder()) e->fused = 1;            // the pipelined kernels are fed by tensor-map TMA
    e->d = d;
    e->device = device;
    e->num_it = d.num_it;
    e->cin0 = 4 * d.num_rx_ant + 2;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) e->num_sms = prop.multiProcessorCount;

    int rc = NRX_OK;
    auto bail = [&](int code) { nrx_destroy(e); return code; };

    // ---- StateInit stacks: per stack 3 sep-convs (9 arrays) ------------------------------------
    const int widths_i[4] = {e->cin0, d.units_init[0], d.units_init[1], d.d_s};
    e->init_layers.resize(3);
    for (int l = 0; l < 3; ++l) {
        const int kpad = l == 0 ? 32 : 128, npad = l == 2 ? 64 : 128;
        rc = build_sep_layer(e->init_layers[l], d.n_io, weight_arrays, weight_sizes, 3 * l, 9, widths_i[l],
                             widths_i[l + 1], kpad, npad, identity_map(widths_i[l]));
        if (rc) return bail(rc);
    }
    for (int m = 0; m < d.n_io; ++m)
        for (int l = 0; l < 3; ++l) e->mac_fixed[m] += int64_t(9 + widths_i[l + 1]) * widths_i[l];
    {
        using LI = StackSmem<kStackInit>;
        std::vector<uint8_t> host(size_t(LI::kBlob) * d.n_io, 0);
        for (int m = 0; m < d.n_io; ++m) {
            rc = pack_stack_blob<kStackInit>(host.data() + size_t(m) * LI::kBlob, weight_arrays, weight_sizes, 9 * m,
                                             widths_i, identity_map(widths_i[0]));
            if (rc) return bail(rc);
            // message MLP of iteration 0 (array sizes are validated in the iteration loop below)
            const int a0 = d.n_io * 9;
            if (weight_sizes[a0] == int64_t(d.d_s) * d.units_agg && weight_sizes[a0 + 2] == int64_t(d.units_agg) * d.d_s)
                pack_stack_agg<kStackInit>(host.data() + size_t(m) * LI::kBlob, weight_arrays, a0, d.d_s, d.units_agg);
        }
        for (int m = 0; m < d.n_io; ++m) {
            nrx_engine::StackBias sb{};
            memcpy(sb.v, host.data() + size_t(m) * LI::kBlob + LI::oBias, sizeof sb.v);
            e->init_bias.push_back(sb);
        }
        if (cudaMalloc(&e->stack_init_blob, host.size()) != cudaSuccess ||
            cudaMemcpy(e->stack_init_blob, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(NRX_ERR_CUDA, "uploading StateInit stack weights failed"));
#ifdef NRX_EXPERIMENTAL_PLANS
        using PI = StackPairSmem<kStackInit>;
        std::vector<uint8_t> ph(size_t(PI::kBlob) * d.n_io * 2, 0);
        for (int m = 0; m < d.n_io; ++m)
            for (int r = 0; r < 2; ++r)
                pack_stack_pair_blob<kStackInit>(ph.data() + (size_t(m) * 2 + r) * PI::kBlob, r, weight_arrays, 9 * m, widths_i,
                                                 identity_map(widths_i[0]));
        if (cudaMalloc(&e->pair_init_blob, ph.size()) != cudaSuccess ||
            cudaMemcpy(e->pair_init_blob, ph.data(), ph.size(), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(NRX_ERR_CUDA, "uploading StateInit pair weights failed"));
#endif
    }

    // ---- iterations -----------------------------------------------------------------------------
    int idx = d.n_io * 9;
    std::vector<int> upd_map(2 * d.d_s + 2);
    for (int c = 0; c < 2 * d.d_s + 2; ++c) upd_map[c] = c < d.d_s ? c : 64 + (c - d.d_s);   // [a | s, pe]
    const int widths_u[4] = {2 * d.d_s + 2, d.units_state[0], d.units_state[1], d.d_s};
    e->upd_layers.resize(d.num_it);
    e->agg_blobs.resize(d.num_it, nullptr);
    e->agg_bias.resize(d.num_it);
    e->stack_upd_blobs.resize(d.num_it, nullptr);
#ifdef NRX_EXPERIMENTAL_PLANS
    e->pair_upd_blobs.resize(d.num_it, nullptr);
#endif
    for (int it = 0; it < d.num_it; ++it) {
        if (weight_sizes[idx] != int64_t(d.d_s) * d.units_agg || weight_sizes[idx + 1] != d.units_agg ||
            weight_sizes[idx + 2] != int64_t(d.units_agg) * d.d_s || weight_sizes[idx + 3] != d.d_s)
            return bail(fail(NRX_ERR_INVALID, "AggregateUserStates Dense layers expected at weight index %d", idx));
        std::vector<uint8_t> host(kAggBlob, 0);
        pack_pw(host.data(), weight_arrays[idx], d.d_s, d.units_agg, 64, identity_map(d.d_s));
        pack_pw(host.data() + 8192, weight_arrays[idx + 2], d.units_agg, d.d_s, 64, identity_map(d.units_agg));
        float* b1 = reinterpret_cast<float*>(host.data() + 16384);
        for (int n = 0; n < d.units_agg; ++n) b1[n] = weight_arrays[idx + 1][n];
        for (int n = 0; n < d.d_s; ++n) b1[64 + n] = weight_arrays[idx + 3][n];
        std::memcpy(e->agg_bias[it].v, b1, sizeof(e->agg_bias[it].v));
        if (cudaMalloc(&e->agg_blobs[it], kAggBlob) != cudaSuccess ||
            cudaMemcpy(e->agg_blobs[it], host.data(), kAggBlob, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(NRX_ERR_CUDA, "uploading aggregation weights failed"));
        idx += 4;
        {
            using LU = StackSmem<kStackUpdate>;
            std::vector<uint8_t> sb(LU::kBlob, 0);
            rc = pack_stack_blob<kStackUpdate>(sb.data(), weight_arrays, weight_sizes, idx, widths_u, upd_map);
            if (rc) return bail(rc);
#ifdef NRX_EXPERIMENTAL_PLANS
            {   // plan 4: the same taps as half2 per channel pair and the biases, passed as kernel parameters
                nrx_engine::TmConsts tc{};
                const int t_off[3] = {LU::oTap1, LU::oTap2, LU::oTap3};
                for (int l = 0; l < 3; ++l)
                    for (int c = 0; c < 64; ++c)
                        for (int k = 0; k < 9; ++k) memcpy(&tc.tap[l][c][k], sb.data() + t_off[l] + (k * 128 + 2 * c) * 2, 4);
                const float* b_in = reinterpret_cast<const float*>(sb.data() + LU::oBias);
                const int b_off[3] = {0, 128, 256}, n_pad[3] = {128, 128, 64};
                for (int l = 0; l < 3; ++l)                // biases and B-image rows in the fragment's column order
                    for (int n = 0; n < n_pad[l]; ++n) tc.bias[b_off[l] + tm_phys_col(n)] = b_in[b_off[l] + n];
                e->tm_consts.push_back(tc);
                std::vector<uint8_t> tb(TmSmem::kW, 0);
                const int w_off[3] = {0, 32768, 65536};
                for (int l = 0; l < 3; ++l) {
                    const int cin = widths_u[l], cout = widths_u[l + 1];
                    const float* w = weight_arrays[idx + 3 * l + 1];
                    for (int c = 0; c < cin; ++c)
                        for (int n = 0; n < cout; ++n) {
                            const int k = l == 0 ? upd_map[c] : c;
                            *reinterpret_cast<__half*>(tb.data() + w_off[l] + size_t(k / 64) * n_pad[l] * 128 +
                                                       sw128_offset(tm_phys_col(n), k % 64)) = __float2half(w[size_t(c) * cout + n]);
                        }
                }
                uint8_t* dev = nullptr;
                if (cudaMalloc(&dev, tb.size()) != cudaSuccess ||
                    cudaMemcpy(dev, tb.data(), tb.size(), cudaMemcpyHostToDevice) != cudaSuccess)
                    return bail(fail(NRX_ERR_CUDA, "uploading plan-4 stack weights failed"));
                e->tm_blobs.push_back(dev);
            }
#endif
            if (it + 1 < d.num_it) {     // message MLP of the next iteration (4 agg + 9 sep-conv arrays per iteration)
                const int an = idx + 9;
                if (weight_sizes[an] != int64_t(d.d_s) * d.units_agg || weight_sizes[an + 2] != int64_t(d.units_agg) * d.d_s)
                    return bail(fail(NRX_ERR_INVALID, "AggregateUserStates Dense layers expected at weight index %d", an));
                pack_stack_agg<kStackUpdate>(sb.data(), weight_arrays, an, d.d_s, d.units_agg);
            }
            {
                nrx_engine::StackBias ub{};
                memcpy(ub.v, sb.data() + LU::oBias, sizeof ub.v);
                e->upd_bias.push_back(ub);
            }
            if (cudaMalloc(&e->stack_upd_blobs[it], sb.size()) != cudaSuccess ||
                cudaMemcpy(e->stack_upd_blobs[it], sb.data(), sb.size(), cudaMemcpyHostToDevice) != cudaSuccess)
                return bail(fail(NRX_ERR_CUDA, "uploading UpdateState stack weights failed"));
#ifdef NRX_EXPERIMENTAL_PLANS
            using PU = StackPairSmem<kStackUpdate>;
            std::vector<uint8_t> pb(size_t(PU::kBlob) * 2, 0);
            for (int r = 0; r < 2; ++r)
                pack_stack_pair_blob<kStackUpdate>(pb.data() + size_t(r) * PU::kBlob, r, weight_arrays, idx, widths_u, upd_map);
            if (cudaMalloc(&e->pair_upd_blobs[it], pb.size()) != cudaSuccess ||
                cudaMemcpy(e->pair_upd_blobs[it], pb.data(), pb.size(), cudaMemcpyHostToDevice) != cudaSuccess)
                return bail(fail(NRX_ERR_CUDA, "uploading UpdateState pair weights failed"));
#endif
        }
        e->upd_layers[it].resize(3);
        for (int l = 0; l < 3; ++l) {
            const int npad = l == 2 ? 64 : 128;
            rc = build_sep_layer(e->upd_layers[it][l], 1, weight_arrays, weight_sizes, idx, 0, widths_u[l],
                                 widths_u[l + 1], 128, npad, l == 0 ? upd_map : identity_map(widths_u[l]));
            if (rc) return bail(rc);
            idx += 3;
        }
    }
    e->mac_per_it = 2 * int64_t(d.d_s) * d.units_agg;
    for (int l = 0; l < 3; ++l) e->mac_per_it += int64_t(9 + widths_u[l + 1]) * widths_u[l];

    // ---- read-outs: n_io LLR heads then the channel-estimate head ------------------------------
    {
        const int ch = idx + 4 * d.n_io, n2 = 2 * d.num_rx_ant;
        if (weight_sizes[ch] != int64_t(d.d_s) * d.units_readout || weight_sizes[ch + 1] != d.units_readout ||
            weight_sizes[ch + 2] != int64_t(d.units_readout) * n2 || weight_sizes[ch + 3] != n2)
            return bail(fail(NRX_ERR_INVALID, "ReadoutChEst Dense layers expected at weight index %d", ch));
        std::vector<uint8_t> host(size_t(kRoBlob) * d.n_io, 0);
        std::vector<int> k_llr = identity_map(d.units_readout), k_h(d.units_readout);
        for (int c = 0; c < d.units_readout; ++c) k_h[c] = 128 + c;
        for (int m = 0; m < d.n_io; ++m) {
            const int i = idx + 4 * m, bits = d.io_bits[m];
            if (weight_sizes[i] != int64_t(d.d_s) * d.units_readout || weight_sizes[i + 1] != d.units_readout ||
                weight_sizes[i + 2] != int64_t(d.units_readout) * bits || weight_sizes[i + 3] != bits)
                return bail(fail(NRX_ERR_INVALID, "ReadoutLLRs Dense layers expected at weight index %d", i));
            uint8_t* b = host.data() + size_t(m) * kRoBlob;
            pack_pw(b, weight_arrays[i], d.d_s, d.units_readout, 256, identity_map(d.d_s), 0);
            pack_pw(b, weight_arrays[ch], d.d_s, d.units_readout, 256, identity_map(d.d_s), 128);
            pack_pw(b + kRoW1, weight_arrays[i + 2], d.units_readout, bits, 32, k_llr, 0);
            pack_pw(b + kRoW1, weight_arrays[ch + 2], d.units_readout, n2, 32, k_h, 16);
            float* b1 = reinterpret_cast<float*>(b + kRoW1 + kRoW2);
            for (int n = 0; n < d.units_readout; ++n) {
                b1[n] = weight_arrays[i + 1][n];
                b1[128 + n] = weight_arrays[ch + 1][n];
            }
            for (int n = 0; n < 256; ++n) {   // bias as two fp16 rows of the W1 image (kRoBiasK)
                const __half hi = __float2half(b1[n]), lo = __float2half(b1[n] - __half2float(hi));
                *reinterpret_cast<__half*>(b + sw128_offset(n, kRoBiasK)) = hi;
                *reinterpret_cast<__half*>(b + sw128_offset(n, kRoBiasK + 1)) = lo;
            }
            float* b2 = b1 + 256;
            for (int n = 0; n < bits; ++n) b2[n] = weight_arrays[i + 3][n];
            for (int n = 0; n < n2; ++n) b2[16 + n] = weight_arrays[ch + 3][n];
            e->mac_fixed[m] += int64_t(d.d_s) * d.units_readout * 2 + int64_t(d.units_readout) * (bits + n2);
        }
        if (cudaMalloc(&e->readout_blob, host.size()) != cudaSuccess ||
            cudaMemcpy(e->readout_blob, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(NRX_ERR_CUDA, "uploading read-out weights failed"));
    }

    // ---- geometry tables ----------------------------------------------------------------------
    const int F = d.num_subcarriers, U = d.max_num_tx, TF = kT * F;
    e->n_pilot_slots = d.num_dmrs_symbols * F;
    {
        // FOCC / CDM de-spreading table: the estimate of pilot slot k is the sum over the non-zero
        // pilots of its block of focc_block consecutive slots of y/p, divided by 2
        // (twin in the reference: utils/neural_rx.py:1620-1629).
        std::vector<FoccEntry> tab(size_t(U) * e->n_pilot_slots);
        for (int u = 0; u < U; ++u)
            for (int k = 0; k < e->n_pilot_slots; ++k) {
                FoccEntry en{};
                const float* pk = pilots + (size_t(u) * e->n_pilot_slots + k) * 2;
                if (pk[0] != 0.f || pk[1] != 0.f) {
                    const int b0 = k / d.focc_block * d.focc_block;
                    int n = 0;
                    for (int j = b0; j < b0 + d.focc_block; ++j) {
                        const float* pj = pilots + (size_t(u) * e->n_pilot_slots + j) * 2;
                        const float mag = pj[0] * pj[0] + pj[1] * pj[1];
                        if (mag == 0.f) continue;
                        if (n >= 2) return bail(fail(NRX_ERR_UNSUPPORTED, "more than two pilots per FOCC block"));
                        en.src[n] = (d.dmrs_symbols[j / F] << 16) | (j % F);
                        en.w[n] = make_float2(0.5f * pj[0] / mag, -0.5f * pj[1] / mag);   // 0.5 / p
                        ++n;
                    }
                }
                tab[size_t(u) * e->n_pilot_slots + k] = en;
            }
        // resolve the nearest-pilot gather now: one entry per (user, RE), (t, f)-major like y
        std::vector<FoccEntry> per_re(size_t(U) * TF);
        for (int u = 0; u < U; ++u)
            for (int f = 0; f < F; ++f)
                for (int t = 0; t < kT; ++t) {
                    const int k = nn_index[size_t(u) * TF + t * F + f];
                    if (k < 0 || k >= e->n_pilot_slots) return bail(fail(NRX_ERR_INVALID, "nn_index out of range"));
                    per_re[size_t(u) * TF + t * F + f] = tab[size_t(u) * e->n_pilot_slots + k];
                }
        if (cudaMalloc(&e->focc, per_re.size() * sizeof(FoccEntry)) != cudaSuccess ||
            cudaMemcpy(e->focc, per_re.data(), per_re.size() * sizeof(FoccEntry), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMalloc(&e->nn_index, size_t(U) * TF * 4) != cudaSuccess ||
            cudaMemcpy(e->nn_index, nn_index, size_t(U) * TF * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMalloc(&e->pos_enc, size_t(U) * TF * 2 * 4) != cudaSuccess ||
            cudaMemcpy(e->pos_enc, pos_enc, size_t(U) * TF * 2 * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMalloc(&e->data_index, size_t(TF) * 4) != cudaSuccess ||
            cudaMemcpy(e->data_index, data_index, size_t(TF) * 4, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(NRX_ERR_CUDA, "uploading geometry tables failed"));
    }

    // ---- kernel attributes ---------------------------------------------------------------------
    cudaError_t ce = cudaSuccess;
    auto acc = [&](cudaError_t r) { if (ce == cudaSuccess) ce = r; };
    acc(set_smem(nrx_sepconv_kernel<32, 128, kHidden>, SepSmem<32, 128>::kTotal));
    acc(set_smem(nrx_sepconv_kernel<128, 128, kHidden>, SepSmem<128, 128>::kTotal));
    acc(set_smem(nrx_sepconv_kernel<128, 64, kInitOut>, SepSmem<128, 64>::kTotal));
    acc(set_smem(nrx_sepconv_kernel<128, 64, kUpdateOut>, SepSmem<128, 64>::kTotal));
    acc(set_smem(nrx_agg_kernel<1>, agg_smem_bytes(1)));
    acc(set_smem(nrx_agg_kernel<2>, agg_smem_bytes(2)));
    acc(set_smem(nrx_agg_kernel<3>, agg_smem_bytes(3)));
    acc(set_smem(nrx_agg_kernel<4>, agg_smem_bytes(4)));
    acc(set_smem(nrx_agg_ws_kernel, kAggWsSmem));
    acc(set_smem(nrx_readout_kernel, kRoSmem));
    acc(set_smem(nrx_stack_kernel<kStackInit, false>, StackSmem<kStackInit>::kTotal));
    acc(set_smem(nrx_stack_kernel<kStackUpdate, false>, StackSmem<kStackUpdate>::kTotal));
#ifdef NRX_EXPERIMENTAL_PLANS
    acc(set_smem(nrx_stack_pair_kernel<kStackInit>, StackPairSmem<kStackInit>::kTotal));
    acc(set_smem(nrx_stack_pair_kernel<kStackUpdate>, StackPairSmem<kStackUpdate>::kTotal));
#endif
    acc(set_smem(nrx_stack_kernel<kStackInit, true>, StackSmem<kStackInit>::kTotal));
    acc(set_smem(nrx_stack_kernel<kStackUpdate, true>, StackSmem<kStackUpdate>::kTotal));
#ifdef NRX_EXPERIMENTAL_PLANS
    acc(set_smem(nrx_stack_tm_kernel, TmSmem::kTotal));
#endif
    acc(set_smem(nrx_stack_ws_kernel<kStackInit>, WsSmem<kStackInit>::kTotal));
    acc(set_smem(nrx_stack_ws_kernel<kStackUpdate>, WsSmem<kStackUpdate>::kTotal));
    if (ce != cudaSuccess) return bail(fail(NRX_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(ce)));
    ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) return bail(fail(NRX_ERR_CUDA, "cudaDeviceSynchronize failed: %s", cudaGetErrorString(ce)));
    *out = e;
    return NRX_OK;
}

int nrx_set_num_it(nrx_engine* e, int32_t num_it) {
    if (!e) return fail(NRX_ERR_INVALID, "null engine");
    if (num_it < 1 || num_it > e->d.num_it) return fail(NRX_ERR_INVALID, "Invalid number of iterations");
    e->num_it = num_it;
    return NRX_OK;
}

int nrx_get_num_it(const nrx_engine* e, int32_t* num_it) {
    if (!e || !num_it) return fail(NRX_ERR_INVALID, "null argument");
    *num_it = e->num_it;
    return NRX_OK;
}

int nrx_set_profiling(nrx_engine* e, int32_t enable) {
    if (!e) return fail(NRX_ERR_INVALID, "null engine");
    e->profiling = enable != 0;
    return NRX_OK;
}

int nrx_get_profile(nrx_engine* e, double* ms, int64_t* launches) {
    if (!e || !ms || !launches) return fail(NRX_ERR_INVALID, "nrx_get_profile: null argument");
    NRX_CUDA(cudaSetDevice(e->device));
    for (auto& s : e->spans) {
        NRX_CUDA(cudaEventSynchronize(s.b));
        float t = 0.f;
        NRX_CUDA(cudaEventElapsedTime(&t, s.a, s.b));
        ms[s.cls] += t;
        launches[s.cls] += 1;
        e->event_pool.push_back(s.a);
        e->event_pool.push_back(s.b);
    }
    e->spans.clear();
    return NRX_OK;
}

int nrx_set_fused(nrx_engine* e, int32_t fused) {
    if (!e) return fail(NRX_ERR_INVALID, "null engine");
    if (fused < 0 || fused > 6) return fail(NRX_ERR_INVALID, "fused must be 0 ... 6");
#ifndef NRX_EXPERIMENTAL_PLANS
    if (fused == 3 || fused == 4)
        return fail(NRX_ERR_UNSUPPORTED, "plans 3 and 4 are experiments that are not in the default build (-DNRX_EXPERIMENTAL_PLANS)");
#endif
    if (fused >= 4 && !tensor_map_encoder()) return fail(NRX_ERR_CUDA, "plans 4, 5 and 6 need cuTensorMapEncodeTiled (driver entry point not found)");
    e->fused = fused;
    return NRX_OK;
}

int nrx_set_skip_inactive(nrx_engine* e, int32_t enable) {
    if (!e) return fail(NRX_ERR_INVALID, "null engine");
    e->skip_inactive = enable != 0;
    return NRX_OK;
}

int nrx_debug_option(nrx_engine* e, int32_t option, int32_t value) {
    if (!e) return fail(NRX_ERR_INVALID, "null engine");
    if (option == NRX_OPT_AGG_PIPELINED) {
        e->agg_pipelined = value != 0;
        return NRX_OK;
    }
    if (option == NRX_OPT_STACK_BALANCED) {
        e->stack_balanced = value != 0;
        return NRX_OK;
    }
    return fail(NRX_ERR_INVALID, "nrx_debug_option: unknown option %d", int(option));
}

int nrx_set_slots_per_pass(nrx_engine* e, int32_t slots) {
    if (!e || slots < 0) return fail(NRX_ERR_INVALID, "slots_per_pass must be >= 0");
    e->slots_per_pass = slots;
    return NRX_OK;
}

int nrx_workspace_bytes(const nrx_engine* e, int32_t batch, size_t* bytes) {
    if (!e || !bytes || batch < 1) return fail(NRX_ERR_INVALID, "nrx_workspace_bytes: bad argument");
    *bytes = layout(e, batch).total;
    return NRX_OK;
}

int nrx_launches_per_forward(const nrx_engine* e, int32_t batch, int32_t* launches) {
    if (!e || !launches || batch < 1) return fail(NRX_ERR_INVALID, "nrx_launches_per_forward: bad argument");
    const int bp = pass_slots(e, batch);
    const int passes = (batch + bp - 1) / bp;
    *launches = e->fused ? 1 + passes * ((e->skip_inactive && (e->fused == 1 || e->fused >= 5) ? 1 : 0) + 1 + 1 + e->num_it * (e->d.max_num_tx == 2 && e->fused == 2 ? 1 : 2) + 1)
                         : 1 + passes * (1 + 3 + e->num_it * 4 + 1);
    return NRX_OK;
}

// Host-only planning helpers (no device needed): how the stack kernels cut the subcarrier axis.
int nrx_plan_stack_chunks(int32_t planes, int32_t num_subcarriers, int32_t num_sms, int32_t* chunks_per_plane) {
    if (planes < 1 || num_subcarriers < 1 || num_sms < 1 || !chunks_per_plane) return fail(NRX_ERR_INVALID, "nrx_plan_stack_chunks: bad argument");
    *chunks_per_plane = choose_chunks(planes, num_subcarriers, num_sms);
    return NRX_OK;
}
int nrx_plan_stack_range(int32_t planes, int32_t num_subcarriers, int32_t num_ctas, int32_t cta, int64_t* first, int64_t* last) {
    if (planes < 1 || num_subcarriers < 1 || num_ctas < 1 || cta < 0 || cta >= num_ctas || !first || !last)
        return fail(NRX_ERR_INVALID, "nrx_plan_stack_range: bad argument");
    long long g0, g1;
    stack_balanced_range((long long)planes * num_subcarriers, num_ctas, cta, g0, g1);
    *first = g0;
    *last = g1;
    return NRX_OK;
}
#ifdef NRX_EXPERIMENTAL_PLANS
int nrx_plan_stack_jobs(int32_t planes, int32_t num_subcarriers, int32_t num_sms, int32_t* jobs_per_plane, int32_t* num_items,
                        int32_t* steps_per_item) {
    if (planes < 1 || num_subcarriers < 1 || num_sms < 1 || !jobs_per_plane || !num_items || !steps_per_item)
        return fail(NRX_ERR_INVALID, "nrx_plan_stack_jobs: bad argument");
    const int j = tm_choose_jobs(planes, num_subcarriers, num_sms);
    *jobs_per_plane = j;
    *num_items = (planes * j + kTmSeqs - 1) / kTmSeqs;
    *steps_per_item = (num_subcarriers + j - 1) / j + kTmFill;
    return NRX_OK;
}
int nrx_fragment_column(int32_t channel) { return channel < 0 ? -1 : tm_phys_col(channel); }

#endif

int nrx_mac_per_pixel(const nrx_engine* e, int32_t llr_head, int64_t* macs) {
    if (!e || !macs || llr_head < 0 || llr_head >= e->d.n_io) return fail(NRX_ERR_INVALID, "nrx_mac_per_pixel: bad argument");
    *macs = e->mac_fixed[llr_head] + e->mac_per_it * e->num_it;
    return NRX_OK;
}

}  // extern "C"

namespace {

struct AerialIn {             // inputs of the Aerial-shaped call (null for the Sionna-shaped one)
    const float *y_re, *y_im, *h_re, *h_im;
};

int forward_impl(nrx_engine* e, void* cuda_stream, int32_t batch, const AerialIn* aer, const void* y, const float* active_tx,
                 const int32_t* io_index, const int32_t* head_index, int32_t llr_head, int32_t out_bits, float* llr,
                 float* llr_grid, float* llr_aerial, float* h_hat_refined, float* h_hat_ls, void* workspace,
                 size_t workspace_bytes) {
    if (!e || (!y && !aer) || !active_tx || !workspace) return fail(NRX_ERR_INVALID, "nrx_forward: null argument");
    if (batch < 1) return fail(NRX_ERR_INVALID, "batch must be >= 1");
    const nrx_model_desc& d = e->d;
    if (llr_head < 0 || llr_head >= d.n_io) return fail(NRX_ERR_INVALID, "llr_head out of range");
    if (out_bits < 1 || out_bits > 16) return fail(NRX_ERR_INVALID, "out_bits out of range");
    const Workspace w = layout(e, batch);
    if (workspace_bytes < w.total)
        return fail(NRX_ERR_WORKSPACE, "workspace has %zu bytes, %zu needed", workspace_bytes, w.total);
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return fail(NRX_ERR_INVALID, "workspace must be 256-byte aligned");
    NRX_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    float* partial = reinterpret_cast<float*>(ws + w.partial);
    __half* z0 = reinterpret_cast<__half*>(ws + w.z0);
    __half* h1 = reinterpret_cast<__half*>(ws + w.h1);
    __half* h2 = reinterpret_cast<__half*>(ws + w.h2);
    __half* abuf = reinterpret_cast<__half*>(ws + w.abuf);
    __half* sbuf = reinterpret_cast<__half*>(ws + w.sbuf);

    const int F = d.num_subcarriers, U = d.max_num_tx, N = d.num_rx_ant;
    const int per_slot = F * kT;
    const float* pe_tab = aer ? e->pos_enc_aerial : e->pos_enc;
    {
        Timed t(e, st, NRX_K_POWER);
        if (aer) nrx_power_planar_kernel<<<dim3(kPowerParts, batch), 256, 0, st>>>(aer->y_re, aer->y_im, partial, N * per_slot);
        else nrx_power_kernel<<<dim3(kPowerParts, batch), 256, 0, st>>>(static_cast<const float2*>(y), partial, N * per_slot);
    }

    const int bp_max = pass_slots(e, batch);
    for (int b0 = 0; b0 < batch; b0 += bp_max) {
        const int bp = batch - b0 < bp_max ? batch - b0 : bp_max;
        const int BU = bp * U;
        PrepParams pp{};
        pp.y = static_cast<const float2*>(y);
        pp.partial = partial;
        pp.focc_re = e->focc;
        pp.pos_enc = e->pos_enc;
        pp.z0 = z0;
        pp.h_ls = h_hat_ls;
        pp.F = F; pp.U = U; pp.N = N; pp.n_pilot_slots = e->n_pilot_slots; pp.b0 = b0; pp.bp = bp;
        if (aer) {
            PrepAerialParams pa{};
            pa.y_re = aer->y_re; pa.y_im = aer->y_im; pa.h_re = aer->h_re; pa.h_im = aer->h_im;
            pa.partial = partial; pa.nn_prb = e->nn_prb; pa.pos_enc = pe_tab; pa.z0 = z0;
            pa.F = F; pa.U = U; pa.N = N; pa.n_pilots = e->aerial_pilots; pa.b0 = b0; pa.bp = bp;
            Timed t(e, st, NRX_K_PREP);
            if (N == 4) nrx_prep_aerial_kernel<4><<<(bp * per_slot + 255) / 256, 256, 0, st>>>(pa);
            else if (N == 2) nrx_prep_aerial_kernel<2><<<(bp * per_slot + 255) / 256, 256, 0, st>>>(pa);
            else nrx_prep_aerial_kernel<0><<<(bp * per_slot + 255) / 256, 256, 0, st>>>(pa);
        } else {
            Timed t(e, st, NRX_K_PREP);
            const int pgrid_ = bp * ((F + kPrepF - 1) / kPrepF);
            if (N == 4) nrx_prep_kernel<4><<<pgrid_, 256, 0, st>>>(pp);
            else if (N == 2) nrx_prep_kernel<2><<<pgrid_, 256, 0, st>>>(pp);
            else nrx_prep_kernel<0><<<pgrid_, 256, 0, st>>>(pp);
        }

        // inactive-user skipping: ordered list of the active (slot, user) planes of this pass, built on the device
        const int skip = e->skip_inactive && e->fused != 0 && e->fused != 2 && e->fused != 3 && e->fused != 4;
        int32_t* plist = skip ? reinterpret_cast<int32_t*>(ws + w.plist) : nullptr;
        if (skip) nrx_planes_kernel<<<1, 1024, 0, st>>>(active_tx + size_t(b0) * U, BU, plist);
        __half* s_cur = sbuf;
        if (e->fused) {
            // ---- fused stacks: StateInit, then per iteration aggregation + fused UpdateState ----
            __half* s_alt = reinterpret_cast<__half*>(ws + w.sbuf2);
            StackParams kp{};
            kp.F = F; kp.U = U; kp.d_s = d.d_s;
            kp.n_stacks = d.n_io;
            kp.plane_list = plist;
            const int sgrid = set_stack_work(e, kp, BU, F);
            kp.pos_enc = pe_tab;
            // two users: the message MLP of the next AggregateUserStates runs in the tail of each stack
            // and user u reads the other user's sp tensor directly (no aggregation kernel, no `a` tensor)
            const bool pair = U == 2 && e->fused == 2;
#ifdef NRX_EXPERIMENTAL_PLANS
            // plan 3: CTA-pair stack kernels (two users, one stack for both users of a slot)
            const bool cta_pair = U == 2 && e->fused == 3;
            StackParams kq = kp;                        // geometry of the pair launches: items = (slot, chunk)
            kq.n_chunks = choose_chunks(bp, F, e->num_sms / 2 > 0 ? e->num_sms / 2 : 1);
            kq.num_items = kq.n_chunks * bp;
            const int pgrid = 2 * (kq.num_items < e->num_sms / 2 ? kq.num_items : e->num_sms / 2);
#endif
            __half* sp_cur = abuf;
            __half* sp_alt = reinterpret_cast<__half*>(ws + w.abuf2);
            kp.z0 = z0; kp.s_out = s_cur;
            kp.wblob = e->stack_init_blob;
            kp.stack_index = io_index ? io_index + size_t(b0) * U : nullptr;
            kp.default_stack = llr_head;
            kp.active_tx = active_tx + size_t(b0) * U;
            kp.pair_agg = 0;
            kp.sp_out = pair ? sp_cur : nullptr;
            const bool piped = e->fused >= 5;           // plans 5, 6: warp-specialised pipelined UpdateState kernels
            if (e->fused == 5 && !io_index) {           // plan 5 also runs StateInit there (measured slower: 431 vs 401 us)                   // (per-user StateInit stacks, Var-IO: the serial kernel switches weights per item)
                if (const int rc = launch_stack_ws<kStackInit>(e, st, kp, BU, F, e->init_bias[llr_head].v)) return rc;
            } else {
                Timed t(e, st, NRX_K_STACK_INIT);
                if (pair) nrx_stack_kernel<kStackInit, true><<<sgrid, kStackThreads, StackSmem<kStackInit>::kTotal, st>>>(kp);
#ifdef NRX_EXPERIMENTAL_PLANS
                else if (cta_pair && !io_index) {
                    kq.z0 = kp.z0; kq.s_out = kp.s_out; kq.wblob = e->pair_init_blob; kq.stack_index = nullptr;
                    kq.default_stack = llr_head; kq.active_tx = kp.active_tx; kq.pair_agg = 0; kq.sp_out = nullptr;
                    nrx_stack_pair_kernel<kStackInit><<<pgrid, kStackThreads, StackPairSmem<kStackInit>::kTotal, st>>>(kq);
                }
#endif
                else nrx_stack_kernel<kStackInit, false><<<sgrid, kStackThreads, StackSmem<kStackInit>::kTotal, st>>>(kp);
            }
            kp.stack_index = nullptr;
            kp.default_stack = 0;
            kp.n_stacks = 1;
            kp.pair_agg = pair ? 1 : 0;
            for (int it = 0; it < e->num_it; ++it) {
                if (!pair)
                    if (const int rc = launch_agg(e, st, s_cur, abuf, it, active_tx + size_t(b0) * U, U, per_slot, bp, skip)) return rc;
                kp.a_in = pair ? sp_cur : abuf; kp.s_in = s_cur; kp.s_out = s_alt;
                kp.sp_out = pair && it + 1 < e->num_it ? sp_alt : nullptr;
                kp.wblob = e->stack_upd_blobs[it];
#ifdef NRX_EXPERIMENTAL_PLANS
                if (e->fused == 4) {
                    const int rc = launch_stack_tm(e, st, it, kp.a_in, kp.s_in, kp.s_out, BU, F);
                    if (rc) return rc;
                } else
#endif
                if (piped) {
                    if (const int rc = launch_stack_ws<kStackUpdate>(e, st, kp, BU, F, e->upd_bias[it].v)) return rc;
                } else {
                    Timed t(e, st, NRX_K_STACK_UPD);
                    if (pair) nrx_stack_kernel<kStackUpdate, true><<<sgrid, kStackThreads, StackSmem<kStackUpdate>::kTotal, st>>>(kp);
#ifdef NRX_EXPERIMENTAL_PLANS
                    else if (cta_pair) {
                        kq.a_in = kp.a_in; kq.s_in = kp.s_in; kq.s_out = kp.s_out; kq.wblob = e->pair_upd_blobs[it];
                        kq.stack_index = nullptr; kq.default_stack = 0; kq.pair_agg = 0; kq.sp_out = nullptr;
                        nrx_stack_pair_kernel<kStackUpdate><<<pgrid, kStackThreads, StackPairSmem<kStackUpdate>::kTotal, st>>>(kq);
                    }
#endif
                    else nrx_stack_kernel<kStackUpdate, false><<<sgrid, kStackThreads, StackSmem<kStackUpdate>::kTotal, st>>>(kp);
                }
                __half* tmp = s_cur; s_cur = s_alt; s_alt = tmp;
                tmp = sp_cur; sp_cur = sp_alt; sp_alt = tmp;
            }
        } else {
            SepParams sp{};
            sp.F = F; sp.U = U; sp.d_s = d.d_s;
            sp.n_stacks = d.n_io;
            sp.tiles_per_bu = (F + kTileF - 1) / kTileF;
            sp.num_tiles = sp.tiles_per_bu * BU;
            sp.pos_enc = pe_tab;
            // ---- StateInit (:107-132), stack per user = one-hot mcs_ue_mask (:562-569) -------------
            sp.stack_index = io_index ? io_index + size_t(b0) * U : nullptr;
            sp.default_stack = llr_head;
            sp.src0 = z0; sp.src1 = nullptr; sp.C0 = 32; sp.C1 = 0; sp.out = h1;
            sp.wblob = e->init_layers[0].blob; sp.blob_bytes = e->init_layers[0].blob_bytes;
            launch_sep<32, 128, kHidden>(e, st, sp);
            sp.src0 = h1; sp.C0 = 128; sp.out = h2;
            sp.wblob = e->init_layers[1].blob; sp.blob_bytes = e->init_layers[1].blob_bytes;
            launch_sep<128, 128, kHidden>(e, st, sp);
            sp.src0 = h2; sp.out = sbuf;
            sp.wblob = e->init_layers[2].blob; sp.blob_bytes = e->init_layers[2].blob_bytes;
            launch_sep<128, 64, kInitOut>(e, st, sp);
            // ---- CGNN iterations (:576-593) --------------------------------------------------------
            sp.stack_index = nullptr;
            sp.default_stack = 0;
            sp.n_stacks = 1;
            for (int it = 0; it < e->num_it; ++it) {
                if (const int rc = launch_agg(e, st, sbuf, abuf, it, active_tx + size_t(b0) * U, U, per_slot, bp)) return rc;
                const auto& L = e->upd_layers[it];
                sp.src0 = abuf; sp.src1 = sbuf; sp.C0 = 64; sp.C1 = 64; sp.out = h1;
                sp.wblob = L[0].blob; sp.blob_bytes = L[0].blob_bytes;
                launch_sep<128, 128, kHidden>(e, st, sp);
                sp.src0 = h1; sp.src1 = nullptr; sp.C0 = 128; sp.C1 = 0; sp.out = h2;
                sp.wblob = L[1].blob; sp.blob_bytes = L[1].blob_bytes;
                launch_sep<128, 128, kHidden>(e, st, sp);
                sp.src0 = h2; sp.out = sbuf;
                sp.wblob = L[2].blob; sp.blob_bytes = L[2].blob_bytes;
                launch_sep<128, 64, kUpdateOut>(e, st, sp);
            }
        }
        // ---- read-outs + resource-grid demapping (:582-593, :843-858) --------------------------
        ReadoutParams rp{};
        rp.sbuf = s_cur;
        rp.wblob = e->readout_blob;
        rp.head_index = head_index ? head_index + size_t(b0) * U : nullptr;
        rp.data_index = e->data_index;
        const size_t bu0 = size_t(b0) * U;
        rp.llr = llr ? llr + bu0 * d.num_data_res * out_bits : nullptr;
        rp.llr_grid = llr_grid ? llr_grid + bu0 * per_slot * out_bits : nullptr;
        rp.h_ref = h_hat_refined ? h_hat_refined + bu0 * per_slot * 2 * N : nullptr;
        rp.llr_aerial = llr_aerial ? llr_aerial + bu0 * per_slot * out_bits : nullptr;
        rp.F = F; rp.U = U; rp.N2 = 2 * N; rp.out_bits = out_bits; rp.n_data = d.num_data_res;
        rp.rows_per_bu = per_slot;
        rp.tiles_per_bu = (per_slot + 127) / 128;
        rp.num_tiles = rp.tiles_per_bu * BU;
        rp.default_head = llr_head;
        rp.n_heads = d.n_io;
        rp.plane_list = plist;
        if (skip) {                                    // outputs of inactive users: zeros
            if (rp.llr) NRX_CUDA(cudaMemsetAsync(rp.llr, 0, size_t(BU) * d.num_data_res * out_bits * 4, st));
            if (rp.llr_grid) NRX_CUDA(cudaMemsetAsync(rp.llr_grid, 0, size_t(BU) * per_slot * out_bits * 4, st));
            if (rp.h_ref) NRX_CUDA(cudaMemsetAsync(rp.h_ref, 0, size_t(BU) * per_slot * 2 * N * 4, st));
            if (rp.llr_aerial) NRX_CUDA(cudaMemsetAsync(rp.llr_aerial, 0, size_t(BU) * per_slot * out_bits * 4, st));
        }
        rp.vec = ((reinterpret_cast<uintptr_t>(rp.llr) | reinterpret_cast<uintptr_t>(rp.llr_grid) |
                   reinterpret_cast<uintptr_t>(rp.h_ref)) & 15u) == 0;
        const int grid = rp.num_tiles < e->num_sms ? rp.num_tiles : e->num_sms;
        {
            Timed t(e, st, NRX_K_READOUT);
            nrx_readout_kernel<<<grid, kThreads, kRoSmem, st>>>(rp);
        }
    }
    NRX_CUDA(cudaGetLastError());
    return NRX_OK;
}

}  // namespace

extern "C" {

int nrx_forward(nrx_engine* e, void* cuda_stream, int32_t batch, const void* y, const float* active_tx,
                const int32_t* io_index, const int32_t* head_index, int32_t llr_head, int32_t out_bits, float* llr,
                float* llr_grid, float* h_hat_refined, float* h_hat_ls, void* workspace, size_t workspace_bytes) {
    if (!y) return fail(NRX_ERR_INVALID, "nrx_forward: null argument");
    return forward_impl(e, cuda_stream, batch, nullptr, y, active_tx, io_index, head_index, llr_head, out_bits, llr, llr_grid,
                        nullptr, h_hat_refined, h_hat_ls, workspace, workspace_bytes);
}

// NRPreprocessing._calculate_nn_indices (utils/neural_rx.py:1631-1670) on the 12 x T template of one
// PRB, tiled over the PRBs: nearest non-zero pilot per RE (Manhattan distance, candidates enumerated
// subcarrier-major / symbol-minor, first minimum wins) and the positional encoding of the template.
int nrx_set_aerial_dmrs(nrx_engine* e, const int32_t* dmrs_ofdm_pos, int32_t n_sym, const int32_t* dmrs_subcarrier_pos,
                        int32_t n_sc) {
    if (!e || !dmrs_ofdm_pos || !dmrs_subcarrier_pos) return fail(NRX_ERR_INVALID, "nrx_set_aerial_dmrs: null argument");
    const nrx_model_desc& d = e->d;
    const int F = d.num_subcarriers, U = d.max_num_tx, T = kT;
    if (F % 12) return fail(NRX_ERR_INVALID, "the Aerial entry point needs whole PRBs");
    if (n_sym < 1 || n_sym > NRX_MAX_DMRS || n_sc < 2 || n_sc > 12 || n_sc % 2)
        return fail(NRX_ERR_INVALID, "need 1..4 DMRS symbols and an even number (<= 12) of pilots per PRB");
    if (d.n_io != 1) return fail(NRX_ERR_UNSUPPORTED, "NeuralReceiverONNX has no support for mixed MCS");
    for (int i = 0; i < U * n_sym; ++i)
        if (dmrs_ofdm_pos[i] < 0 || dmrs_ofdm_pos[i] >= T) return fail(NRX_ERR_INVALID, "dmrs_ofdm_pos out of range");
    for (int i = 0; i < U * n_sc; ++i)
        if (dmrs_subcarrier_pos[i] < 0 || dmrs_subcarrier_pos[i] >= 12)
            return fail(NRX_ERR_INVALID, "dmrs_subcarrier_pos out of range");
    const int n_prb = F / 12, per_sym = n_prb * n_sc, TF = T * F;
    std::vector<int32_t> nn(size_t(U) * TF);
    std::vector<float> pe(size_t(U) * TF * 2);
    for (int u = 0; u < U; ++u) {
        const int32_t* tp = dmrs_ofdm_pos + u * n_sym;
        const int32_t* sp = dmrs_subcarrier_pos + u * n_sc;
        double dist[2][12 * kT], mean[2] = {0, 0}, var[2] = {0, 0};
        int best_k[12 * kT], best_j[12 * kT];
        for (int sc = 0; sc < 12; ++sc)
            for (int t = 0; t < T; ++t) {
                int best = 1 << 30, bk = 0, bj = 0, dt_min = 1 << 30, df_min = 1 << 30;
                for (int k = 0; k < n_sc; ++k)
                    for (int j = 0; j < n_sym; ++j) {
                        const int df = std::abs(sc - sp[k]), dt = std::abs(t - tp[j]);
                        if (df + dt < best) { best = df + dt; bk = k; bj = j; }
                        if (dt < dt_min) dt_min = dt;
                        if (df < df_min) df_min = df;
                    }
                best_k[sc * T + t] = bk;
                best_j[sc * T + t] = bj;
                dist[0][sc * T + t] = dt_min;
                dist[1][sc * T + t] = df_min;
            }
        for (int c = 0; c < 2; ++c) {
            for (int i = 0; i < 12 * T; ++i) mean[c] += dist[c][i];
            mean[c] /= 12 * T;
            for (int i = 0; i < 12 * T; ++i) var[c] += (dist[c][i] - mean[c]) * (dist[c][i] - mean[c]);
            var[c] = std::sqrt(var[c] / (12 * T)) + 1e-8;                      // population std + 1e-8 (:1655-1658)
        }
        for (int f = 0; f < F; ++f)
            for (int t = 0; t < T; ++t) {
                const int prb = f / 12, i = (f % 12) * T + t;
                nn[size_t(u) * TF + f * T + t] = best_j[i] * per_sym + prb * n_sc + best_k[i];
                for (int c = 0; c < 2; ++c)
                    pe[(size_t(u) * TF + f * T + t) * 2 + c] = float((dist[c][i] - mean[c]) / var[c]);
            }
    }
    NRX_CUDA(cudaSetDevice(e->device));
    if (!e->nn_prb) NRX_CUDA(cudaMalloc(&e->nn_prb, nn.size() * 4));
    if (!e->pos_enc_aerial) NRX_CUDA(cudaMalloc(&e->pos_enc_aerial, pe.size() * 4));
    NRX_CUDA(cudaDeviceSynchronize());
    NRX_CUDA(cudaMemcpy(e->nn_prb, nn.data(), nn.size() * 4, cudaMemcpyHostToDevice));
    NRX_CUDA(cudaMemcpy(e->pos_enc_aerial, pe.data(), pe.size() * 4, cudaMemcpyHostToDevice));
    e->aerial_pilots = n_sym * per_sym;
    return NRX_OK;
}

int nrx_forward_aerial(nrx_engine* e, void* cuda_stream, int32_t batch, const float* rx_slot_real, const float* rx_slot_imag,
                       const float* h_hat_real, const float* h_hat_imag, const float* active_dmrs_ports, float* llr,
                       float* h_hat, void* workspace, size_t workspace_bytes) {
    if (!e || !rx_slot_real || !rx_slot_imag || !h_hat_real || !h_hat_imag)
        return fail(NRX_ERR_INVALID, "nrx_forward_aerial: null argument");
    if (!e->nn_prb) return fail(NRX_ERR_INVALID, "nrx_forward_aerial: call nrx_set_aerial_dmrs first");
    const AerialIn in{rx_slot_real, rx_slot_imag, h_hat_real, h_hat_imag};
    return forward_impl(e, cuda_stream, batch, &in, nullptr, active_dmrs_ports, nullptr, nullptr, 0, e->d.io_bits[0], nullptr,
                        nullptr, llr, h_hat, nullptr, workspace, workspace_bytes);
}

#ifdef NRX_PHASE_TIMING
// debug build only (tools/phase_timing.py): read and reset the stack kernel's per-phase cycle counters
int nrx_debug_phase_cycles(unsigned long long* out32) {
    unsigned long long zero[32] = {0};
    if (cudaMemcpyFromSymbol(out32, g_phase_cycles, sizeof zero) != cudaSuccess) return NRX_ERR_CUDA;
    if (cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof zero) != cudaSuccess) return NRX_ERR_CUDA;
    return NRX_OK;
}
int nrx_debug_ws_cycles(unsigned long long* out48) {
    unsigned long long zero[48] = {0};
    if (cudaMemcpyFromSymbol(out48, g_ws_cycles, sizeof zero) != cudaSuccess) return NRX_ERR_CUDA;
    if (cudaMemcpyToSymbol(g_ws_cycles, zero, sizeof zero) != cudaSuccess) return NRX_ERR_CUDA;
    return NRX_OK;
}
#ifdef NRX_EXPERIMENTAL_PLANS
int nrx_debug_tm_cycles(unsigned long long* out32) {
    unsigned long long zero[32] = {0};
    if (cudaMemcpyFromSymbol(out32, g_tm_cycles, sizeof zero) != cudaSuccess) return NRX_ERR_CUDA;
    if (cudaMemcpyToSymbol(g_tm_cycles, zero, sizeof zero) != cudaSuccess) return NRX_ERR_CUDA;
    return NRX_OK;
}
#endif
#endif

// Asynchronous host-buffer call.  All buffers must be page-locked (cudaHostAlloc / cudaHostRegister / pinned torch
// tensors): they are DMA'd in place.  The call only enqueues: chunk i's H2D copy, kernels and D2H copies go to three
// streams and are ordered by events; the ring of kRing chunk buffers runs on across calls, so the first copy-in of
// call n+1 overlaps the kernels and copy-outs of call n.  Up to kCalls calls may be in flight; the buffers of a
// call belong to the engine until nrx_wait(ticket) has returned.
int nrx_forward_host_async(nrx_engine* e, int32_t batch, const void* y, const float* active_tx, const int32_t* io_index,
                           const int32_t* head_index, int32_t llr_head, int32_t out_bits, float* llr, float* llr_grid,
                           float* h_hat_refined, float* h_hat_ls, int64_t* ticket) {
    if (!e || !y || !active_tx || !ticket) return fail(NRX_ERR_INVALID, "nrx_forward_host_async: null argument");
    if (batch < 1) return fail(NRX_ERR_INVALID, "batch must be >= 1");
    if (out_bits < 1 || out_bits > 16) return fail(NRX_ERR_INVALID, "out_bits out of range");
    const nrx_model_desc& d = e->d;
    NRX_CUDA(cudaSetDevice(e->device));
    constexpr int R = nrx_engine::kRing, NC = nrx_engine::kCalls;
    if (int rc = host_streams(e)) return rc;
    auto pinned = [](const void* p) {
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    float* outs[4] = {llr, llr_grid, h_hat_refined, h_hat_ls};
    bool ok = pinned(y) && pinned(active_tx) && (!io_index || pinned(io_index)) && (!head_index || pinned(head_index));
    for (int k = 0; k < 4; ++k) ok = ok && (!outs[k] || pinned(outs[k]));
    if (!ok) return fail(NRX_ERR_INVALID, "nrx_forward_host_async needs page-locked buffers (use nrx_forward_host for pageable memory)");

    const size_t U = d.max_num_tx, per_slot = size_t(d.num_subcarriers) * kT, N2 = 2 * d.num_rx_ant;
    // chunking: with calls overlapping each other there is no copy to hide INSIDE a call, and the kernels run best on
    // large batches: whole calls up to 32 slots (measured on B200, 30-slot steps: 0.985 of the device-resident
    // throughput with one chunk per call, 0.93 with three)
    int c = e->host_chunk > 0 ? e->host_chunk : 32;
    if (c > batch) c = batch;
    const size_t y_slot = size_t(d.num_rx_ant) * per_slot * 8;
    // ring buffers are sized for the widest output set (16 values per RE) so that calls with different heads share them
    const size_t out_cap[4] = {U * d.num_data_res * 16 * 4, U * per_slot * 16 * 4, U * per_slot * N2 * 4, U * per_slot * N2 * 4};
    const size_t out_slot[4] = {U * d.num_data_res * out_bits * 4, U * per_slot * out_bits * 4, U * per_slot * N2 * 4,
                                U * per_slot * N2 * 4};
    int mask = 0;
    for (int k = 0; k < 4; ++k) mask |= outs[k] ? 1 << k : 0;
    const size_t small_cap = align_up(size_t(4096) * U * 4, 256);          // per call: active_tx | io_index | head_index
    if (size_t(batch) * U * 4 > small_cap) return fail(NRX_ERR_INVALID, "batch too large for the asynchronous call (<= 4096 slots)");
    // device arena: [NC x 3 small arrays][R x (y chunk | 4 output chunks)]
    size_t off = 0;
    auto piece = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_small[NC][3];
    for (int q = 0; q < NC; ++q)
        for (int k = 0; k < 3; ++k) o_small[q][k] = piece(small_cap);
    const size_t C = size_t(c) > e->arena_chunk_slots ? size_t(c) : e->arena_chunk_slots;
    const int amask = mask | e->arena_outs;
    size_t o_y[R], o_out[R][4];
    for (int r = 0; r < R; ++r) {
        o_y[r] = piece(C * y_slot);
        for (int k = 0; k < 4; ++k) o_out[r][k] = piece((amask >> k) & 1 ? C * out_cap[k] : 0);
    }
    if (off > e->d_io_bytes || C != e->arena_chunk_slots || amask != e->arena_outs) {     // (re)size: nothing may be in flight
        NRX_CUDA(cudaDeviceSynchronize());
        if (off > e->d_io_bytes) {
            cudaFree(e->d_io);
            e->d_io = nullptr;
            e->d_io_bytes = 0;
            NRX_CUDA(cudaMalloc(&e->d_io, off));
            e->d_io_bytes = off;
        }
        e->arena_chunk_slots = C;
        e->arena_outs = amask;
    }
    const Workspace w = layout(e, int(C));
    if (w.total > e->d_ws_bytes) {
        NRX_CUDA(cudaDeviceSynchronize());
        cudaFree(e->d_ws);
        e->d_ws = nullptr;
        e->d_ws_bytes = 0;
        NRX_CUDA(cudaMalloc(&e->d_ws, w.total));
        e->d_ws_bytes = w.total;
    }
    const int q = int(e->call_seq % NC);
    if (e->call_seq >= uint64_t(NC)) NRX_CUDA(cudaEventSynchronize(e->ev_call[q]));      // the call that used this ticket slot
    uint8_t* dp = static_cast<uint8_t*>(e->d_io);
    const size_t BU = size_t(batch) * U;
    NRX_CUDA(cudaMemcpyAsync(dp + o_small[q][0], active_tx, BU * 4, cudaMemcpyHostToDevice, e->s_h2d));
    if (io_index) NRX_CUDA(cudaMemcpyAsync(dp + o_small[q][1], io_index, BU * 4, cudaMemcpyHostToDevice, e->s_h2d));
    if (head_index) NRX_CUDA(cudaMemcpyAsync(dp + o_small[q][2], head_index, BU * 4, cudaMemcpyHostToDevice, e->s_h2d));
    for (int b0 = 0; b0 < batch; b0 += c) {
        const int n = batch - b0 < c ? batch - b0 : c;
        const int r = int(e->chunk_seq % R);
        if (e->chunk_seq >= uint64_t(R)) NRX_CUDA(cudaStreamWaitEvent(e->s_h2d, e->ev_d2h[r], 0));   // slot's previous chunk is out
        NRX_CUDA(cudaMemcpyAsync(dp + o_y[r], static_cast<const uint8_t*>(y) + size_t(b0) * y_slot, size_t(n) * y_slot,
                                 cudaMemcpyHostToDevice, e->s_h2d));
        NRX_CUDA(cudaEventRecord(e->ev_h2d[r], e->s_h2d));
        NRX_CUDA(cudaStreamWaitEvent(e->stream, e->ev_h2d[r], 0));
        const int rc = nrx_forward(e, e->stream, n, dp + o_y[r], reinterpret_cast<const float*>(dp + o_small[q][0]) + size_t(b0) * U,
                                   io_index ? reinterpret_cast<const int32_t*>(dp + o_small[q][1]) + size_t(b0) * U : nullptr,
                                   head_index ? reinterpret_cast<const int32_t*>(dp + o_small[q][2]) + size_t(b0) * U : nullptr,
                                   llr_head, out_bits, outs[0] ? reinterpret_cast<float*>(dp + o_out[r][0]) : nullptr,
                                   outs[1] ? reinterpret_cast<float*>(dp + o_out[r][1]) : nullptr,
                                   outs[2] ? reinterpret_cast<float*>(dp + o_out[r][2]) : nullptr,
                                   outs[3] ? reinterpret_cast<float*>(dp + o_out[r][3]) : nullptr, e->d_ws, e->d_ws_bytes);
        if (rc) return rc;
        NRX_CUDA(cudaEventRecord(e->ev_comp[r], e->stream));
        NRX_CUDA(cudaStreamWaitEvent(e->s_d2h, e->ev_comp[r], 0));
        for (int k = 0; k < 4; ++k)
            if (outs[k])
                NRX_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(outs[k]) + size_t(b0) * out_slot[k], dp + o_out[r][k],
                                         size_t(n) * out_slot[k], cudaMemcpyDeviceToHost, e->s_d2h));
        NRX_CUDA(cudaEventRecord(e->ev_d2h[r], e->s_d2h));
        ++e->chunk_seq;
    }
    NRX_CUDA(cudaEventRecord(e->ev_call[q], e->s_d2h));
    *ticket = int64_t(e->call_seq++);
    return NRX_OK;
}

// Blocks until the call that returned `ticket` has delivered all its outputs.
int nrx_wait(nrx_engine* e, int64_t ticket) {
    if (!e || ticket < 0 || uint64_t(ticket) >= e->call_seq) return fail(NRX_ERR_INVALID, "nrx_wait: unknown ticket");
    if (e->call_seq - uint64_t(ticket) > uint64_t(nrx_engine::kCalls)) return NRX_OK;      // its slot has been recycled: long done
    NRX_CUDA(cudaSetDevice(e->device));
    NRX_CUDA(cudaEventSynchronize(e->ev_call[ticket % nrx_engine::kCalls]));
    return NRX_OK;
}

// ---- test hooks: ONE kernel of the path on caller-provided device tensors ----------------------------------------
// (tests/test_gpu_kernels.py drives them with random tensors at edge-case widths; not used by the product path)
int nrx_debug_aggregate(nrx_engine* e, void* cuda_stream, int32_t it, int32_t batch, const void* s_f16, const float* active_tx,
                        void* a_f16) {
    if (!e || !s_f16 || !active_tx || !a_f16 || batch < 1 || it < 0 || it >= e->d.num_it)
        return fail(NRX_ERR_INVALID, "nrx_debug_aggregate: bad argument");
    NRX_CUDA(cudaSetDevice(e->device));
    const int rc = launch_agg(e, static_cast<cudaStream_t>(cuda_stream), static_cast<const __half*>(s_f16), static_cast<__half*>(a_f16),
                              it, active_tx, e->d.max_num_tx, e->d.num_subcarriers * kT, batch, e->skip_inactive ? 1 : 0);
    if (rc) return rc;
    NRX_CUDA(cudaGetLastError());
    return NRX_OK;
}

int nrx_debug_stack(nrx_engine* e, void* cuda_stream, int32_t it, int32_t stack, int32_t batch, const void* z0_f16, const void* a_f16,
                    const void* s_f16, void* s_out_f16) {
    if (!e || !s_out_f16 || batch < 1 || it >= e->d.num_it || stack < 0 || stack >= e->d.n_io)
        return fail(NRX_ERR_INVALID, "nrx_debug_stack: bad argument");
    const bool init = it < 0;
    if (init ? !z0_f16 : (!a_f16 || !s_f16)) return fail(NRX_ERR_INVALID, "nrx_debug_stack: null input");
    if (e->fused != 1 && e->fused < 5) return fail(NRX_ERR_INVALID, "nrx_debug_stack: plans 1, 5 and 6 only");
    NRX_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const nrx_model_desc& d = e->d;
    const int F = d.num_subcarriers, BU = batch * d.max_num_tx;
    StackParams kp{};
    kp.F = F; kp.U = d.max_num_tx; kp.d_s = d.d_s;
    const int sgrid = set_stack_work(e, kp, BU, F);
    kp.pos_enc = e->pos_enc;
    kp.z0 = static_cast<const __half*>(z0_f16);
    kp.a_in = static_cast<const __half*>(a_f16);
    kp.s_in = static_cast<const __half*>(s_f16);
    kp.s_out = static_cast<__half*>(s_out_f16);
    kp.wblob = init ? e->stack_init_blob : e->stack_upd_blobs[it];
    kp.default_stack = init ? stack : 0;
    kp.n_stacks = init ? d.n_io : 1;
    if (e->fused == 5 || (e->fused == 6 && !init)) {
        const int rc = init ? launch_stack_ws<kStackInit>(e, st, kp, BU, F, e->init_bias[stack].v)
                            : launch_stack_ws<kStackUpdate>(e, st, kp, BU, F, e->upd_bias[it].v);
        if (rc) return rc;
    } else if (init) {
        nrx_stack_kernel<kStackInit, false><<<sgrid, kStackThreads, StackSmem<kStackInit>::kTotal, st>>>(kp);
    } else {
        nrx_stack_kernel<kStackUpdate, false><<<sgrid, kStackThreads, StackSmem<kStackUpdate>::kTotal, st>>>(kp);
    }
    NRX_CUDA(cudaGetLastError());
    return NRX_OK;
}

int nrx_debug_readout(nrx_engine* e, void* cuda_stream, int32_t head, int32_t batch, int32_t out_bits, const void* s_f16,
                      float* llr_grid, float* h_hat_refined) {
    if (!e || !s_f16 || batch < 1 || head < 0 || head >= e->d.n_io || out_bits < 1 || out_bits > 16)
        return fail(NRX_ERR_INVALID, "nrx_debug_readout: bad argument");
    NRX_CUDA(cudaSetDevice(e->device));
    const nrx_model_desc& d = e->d;
    ReadoutParams rp{};
    rp.sbuf = static_cast<const __half*>(s_f16);
    rp.wblob = e->readout_blob;
    rp.data_index = e->data_index;
    rp.llr_grid = llr_grid;
    rp.h_ref = h_hat_refined;
    rp.F = d.num_subcarriers; rp.U = d.max_num_tx; rp.N2 = 2 * d.num_rx_ant; rp.out_bits = out_bits; rp.n_data = d.num_data_res;
    rp.rows_per_bu = d.num_subcarriers * kT;
    rp.tiles_per_bu = (rp.rows_per_bu + 127) / 128;
    rp.num_tiles = rp.tiles_per_bu * batch * d.max_num_tx;
    rp.default_head = head;
    rp.n_heads = d.n_io;
    rp.vec = ((reinterpret_cast<uintptr_t>(rp.llr_grid) | reinterpret_cast<uintptr_t>(rp.h_ref)) & 15u) == 0;
    const int grid = rp.num_tiles < e->num_sms ? rp.num_tiles : e->num_sms;
    nrx_readout_kernel<<<grid, kThreads, kRoSmem, static_cast<cudaStream_t>(cuda_stream)>>>(rp);
    NRX_CUDA(cudaGetLastError());
    return NRX_OK;
}

int nrx_set_host_chunk(nrx_engine* e, int32_t slots) {
    if (!e || slots < 0) return fail(NRX_ERR_INVALID, "host chunk must be >= 0");
    e->host_chunk = slots;
    return NRX_OK;
}

}  // extern "C"

namespace {
int host_streams(nrx_engine* e) {                      // three streams + ring events of the host-buffer calls
    if (e->stream) return NRX_OK;
    NRX_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    NRX_CUDA(cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking));
    NRX_CUDA(cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking));
    for (int r = 0; r < nrx_engine::kRing; ++r) {
        NRX_CUDA(cudaEventCreateWithFlags(&e->ev_h2d[r], cudaEventDisableTiming));
        NRX_CUDA(cudaEventCreateWithFlags(&e->ev_comp[r], cudaEventDisableTiming));
        NRX_CUDA(cudaEventCreateWithFlags(&e->ev_d2h[r], cudaEventDisableTiming));
    }
    for (int c = 0; c < nrx_engine::kCalls; ++c) NRX_CUDA(cudaEventCreateWithFlags(&e->ev_call[c], cudaEventDisableTiming));
    return NRX_OK;
}
}  // namespace

extern "C" {

// Host-buffer call: the batch is cut into chunks of `host_chunk` slots that flow through a
// three-stage pipeline on three streams — H2D of chunk i+1, kernels of chunk i and D2H of chunk
// i-1 overlap.  Pinned (page-locked / registered) user buffers are DMA'd directly; pageable ones
// are staged through engine-owned pinned memory chunk by chunk, the memcpy overlapping GPU work.
int nrx_forward_host(nrx_engine* e, int32_t batch, const void* y, const float* active_tx, const int32_t* io_index,
                     const int32_t* head_index, int32_t llr_head, int32_t out_bits, float* llr, float* llr_grid,
                     float* h_hat_refined, float* h_hat_ls) {
    if (!e || !y || !active_tx) return fail(NRX_ERR_INVALID, "nrx_forward_host: null argument");
    if (batch < 1) return fail(NRX_ERR_INVALID, "batch must be >= 1");
    if (out_bits < 1 || out_bits > 16) return fail(NRX_ERR_INVALID, "out_bits out of range");
    const nrx_model_desc& d = e->d;
    NRX_CUDA(cudaSetDevice(e->device));
    constexpr int R = nrx_engine::kRing;
    if (int rc = host_streams(e)) return rc;
    // the synchronous call uses the ring from slot 0 with its own layout: drain whatever asynchronous calls left behind
    if (e->arena_chunk_slots) {
        NRX_CUDA(cudaDeviceSynchronize());
        e->chunk_seq = 0;
        e->arena_chunk_slots = 0;
        e->arena_outs = 0;
    }
    const size_t U = d.max_num_tx, per_slot = size_t(d.num_subcarriers) * kT, N2 = 2 * d.num_rx_ant;
    // Chunk schedule: uniform chunks, by default a third of the batch and at most 16 slots (three chunks in flight).
    // A tapered schedule (small first and last chunk, so that less of the first H2D / last D2H copy is exposed) was
    // measured equal with pinned buffers and slower with pageable ones: the un-hidden cost is per chunk, not per byte.
    std::vector<int> cn;                                // slots per chunk
    {
        int c = e->host_chunk > 0 ? e->host_chunk : (batch + 2) / 3;
        if (e->host_chunk <= 0 && c > 16) c = 16;
        if (c > batch) c = batch;
        for (int b0 = 0; b0 < batch; b0 += c) cn.push_back(batch - b0 < c ? batch - b0 : c);
    }
    const int n_chunks = int(cn.size());
    std::vector<int> cb0(n_chunks, 0);                  // first slot of every chunk
    int C = 0;                                          // largest chunk: sizes the ring buffers
    for (int i = 0; i < n_chunks; ++i) {
        cb0[i] = i ? cb0[i - 1] + cn[i - 1] : 0;
        if (cn[i] > C) C = cn[i];
    }
    // bytes per slot of every stream of data
    const size_t y_slot = size_t(d.num_rx_ant) * per_slot * 8;
    float* outs[4] = {llr, llr_grid, h_hat_refined, h_hat_ls};
    const size_t out_slot[4] = {U * d.num_data_res * out_bits * 4, U * per_slot * out_bits * 4, U * per_slot * N2 * 4,
                                U * per_slot * N2 * 4};
    auto pinned = [](const void* p) {
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    const bool y_pinned = pinned(y);
    bool out_pinned[4];
    for (int k = 0; k < 4; ++k) out_pinned[k] = outs[k] && pinned(outs[k]);

    // device arena: [small per-batch inputs][R x (y chunk | 4 output chunks)]
    size_t off = 0;
    auto piece = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t BU = size_t(batch) * U;
    const size_t o_act = piece(BU * 4), o_io = piece(BU * 4), o_head = piece(BU * 4);
    const size_t small_bytes = off;
    size_t o_y[R], o_out[R][4];
    for (int r = 0; r < R; ++r) {
        o_y[r] = piece(size_t(C) * y_slot);
        for (int k = 0; k < 4; ++k) o_out[r][k] = piece(outs[k] ? size_t(C) * out_slot[k] : 0);
    }
    const size_t total = off;
    if (total > e->d_io_bytes) {
        NRX_CUDA(cudaDeviceSynchronize());
        cudaFree(e->d_io);
        e->d_io = nullptr;
        e->d_io_bytes = 0;
        NRX_CUDA(cudaMalloc(&e->d_io, total));
        e->d_io_bytes = total;
    }
    if (total > e->h_pin_bytes) {                      // pinned staging mirrors the device arena
        NRX_CUDA(cudaDeviceSynchronize());
        if (e->h_pin) cudaFreeHost(e->h_pin);
        e->h_pin = nullptr;
        e->h_pin_bytes = 0;
        NRX_CUDA(cudaMallocHost(&e->h_pin, total));
        e->h_pin_bytes = total;
    }
    const Workspace w = layout(e, C);
    if (w.total > e->d_ws_bytes) {
        NRX_CUDA(cudaDeviceSynchronize());
        cudaFree(e->d_ws);
        e->d_ws = nullptr;
        e->d_ws_bytes = 0;
        NRX_CUDA(cudaMalloc(&e->d_ws, w.total));
        e->d_ws_bytes = w.total;
    }
    uint8_t* hp = static_cast<uint8_t*>(e->h_pin);
    uint8_t* dp = static_cast<uint8_t*>(e->d_io);
    memcpy(hp + o_act, active_tx, BU * 4);
    if (io_index) memcpy(hp + o_io, io_index, BU * 4);
    if (head_index) memcpy(hp + o_head, head_index, BU * 4);
    NRX_CUDA(cudaMemcpyAsync(dp, hp, small_bytes, cudaMemcpyHostToDevice, e->s_h2d));

    auto drain = [&](int i) -> int {                   // chunk i: wait for its D2H, un-stage pageable outputs
        const int r = i % R, b0 = cb0[i], n = cn[i];
        NRX_CUDA(cudaEventSynchronize(e->ev_d2h[r]));
        for (int k = 0; k < 4; ++k)
            if (outs[k] && !out_pinned[k])
                par_memcpy(reinterpret_cast<uint8_t*>(outs[k]) + size_t(b0) * out_slot[k], hp + o_out[r][k], size_t(n) * out_slot[k]);
        return NRX_OK;
    };
    for (int i = 0; i < n_chunks; ++i) {
        const int r = i % R, b0 = cb0[i], n = cn[i];
        if (i >= R) { const int rc = drain(i - R); if (rc) return rc; }
        const uint8_t* ysrc = static_cast<const uint8_t*>(y) + size_t(b0) * y_slot;
        if (!y_pinned) {
            par_memcpy(hp + o_y[r], ysrc, size_t(n) * y_slot);
            ysrc = hp + o_y[r];
        }
        NRX_CUDA(cudaMemcpyAsync(dp + o_y[r], ysrc, size_t(n) * y_slot, cudaMemcpyHostToDevice, e->s_h2d));
        NRX_CUDA(cudaEventRecord(e->ev_h2d[r], e->s_h2d));
        NRX_CUDA(cudaStreamWaitEvent(e->stream, e->ev_h2d[r], 0));
        const int rc = nrx_forward(e, e->stream, n, dp + o_y[r], reinterpret_cast<const float*>(dp + o_act) + size_t(b0) * U,
                                   io_index ? reinterpret_cast<const int32_t*>(dp + o_io) + size_t(b0) * U : nullptr,
                                   head_index ? reinterpret_cast<const int32_t*>(dp + o_head) + size_t(b0) * U : nullptr,
                                   llr_head, out_bits, outs[0] ? reinterpret_cast<float*>(dp + o_out[r][0]) : nullptr,
                                   outs[1] ? reinterpret_cast<float*>(dp + o_out[r][1]) : nullptr,
                                   outs[2] ? reinterpret_cast<float*>(dp + o_out[r][2]) : nullptr,
                                   outs[3] ? reinterpret_cast<float*>(dp + o_out[r][3]) : nullptr, e->d_ws, e->d_ws_bytes);
        if (rc) return rc;
        NRX_CUDA(cudaEventRecord(e->ev_comp[r], e->stream));
        NRX_CUDA(cudaStreamWaitEvent(e->s_d2h, e->ev_comp[r], 0));
        for (int k = 0; k < 4; ++k)
            if (outs[k]) {
                uint8_t* dst = out_pinned[k] ? reinterpret_cast<uint8_t*>(outs[k]) + size_t(b0) * out_slot[k] : hp + o_out[r][k];
                NRX_CUDA(cudaMemcpyAsync(dst, dp + o_out[r][k], size_t(n) * out_slot[k], cudaMemcpyDeviceToHost, e->s_d2h));
            }
        NRX_CUDA(cudaEventRecord(e->ev_d2h[r], e->s_d2h));
    }
    for (int i = n_chunks > R ? n_chunks - R : 0; i < n_chunks; ++i) {
        const int rc = drain(i);
        if (rc) return rc;
    }
    return NRX_OK;
}

}  // extern "C"
