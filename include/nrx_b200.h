/* nrx_b200.h — C ABI of the B200-native neural-receiver engine (libnrx_b200.so).
 *
 * The reference (theshubh007/neural_rx) has no FFI layer: its receiver boundary is the Python
 * object API of utils/neural_rx.py.  Each entry point below states which reference interface it
 * stands in for; the Python mirror classes in neural_rx_b200/receiver.py bind them with ctypes
 * (see INTEGRATION.md).  Plain pointers and sizes only; every function returns an int status
 * (0 = NRX_OK) and never throws; nrx_last_error() gives the message of the last failure on the
 * calling thread.
 *
 * Tensor conventions (all row-major, last index fastest):
 *   y            [B][1][N_rx][T][F]      complex64 (interleaved re,im)  — the `y` of
 *                                         NeuralPUSCHReceiver.forward, utils/neural_rx.py:1544-1603
 *   active_tx    [B][U]                   float32 0/1                   — `active_tx`, same call
 *   llr          [B][U][n_data_res*bits]  float32, llr > 0 <=> bit 1    — CGNNOFDM.forward output,
 *                                         utils/neural_rx.py:843-858,881 (RG-demapped, bit fastest)
 *   llr_grid     [B][U][F][T][bits]       float32                       — CGNN.forward output,
 *                                         utils/neural_rx.py:582-592
 *   h_hat_ref    [B][U][F][T][2*N_rx]     float32 (re | im)             — ReadoutChEst, :593
 *   h_hat_ls     [B][U][F][T][2*N_rx]     float32 (re | im)             — estimate_channel, :1462-1514
 */
#ifndef NRX_B200_H
#define NRX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRX_OK 0
#define NRX_ERR_INVALID 1      /* bad argument / shape mismatch (reference: assert / ValueError)   */
#define NRX_ERR_UNSUPPORTED 2  /* architecture outside what the kernels implement
                                  (reference: NotImplementedError("Unknown layer_type selected."))  */
#define NRX_ERR_CUDA 3         /* CUDA runtime failure; message holds cudaGetErrorString            */
#define NRX_ERR_WORKSPACE 4    /* caller workspace too small                                        */

#define NRX_MAX_IO 4           /* StateInit / ReadoutLLRs stacks (one per MCS in Var-IO mode)       */
#define NRX_MAX_DMRS 4
#define NRX_MAX_TX 4

typedef struct nrx_engine nrx_engine;

/* What CGNN.__init__ / CGNNOFDM.__init__ read from sys_parameters
 * (utils/neural_rx.py:407-530, 638-662) plus the PUSCH grid geometry. */
typedef struct nrx_model_desc {
    int32_t num_rx_ant;          /* N_rx in [1, 7]                                                  */
    int32_t max_num_tx;          /* U in [1, NRX_MAX_TX]                                            */
    int32_t num_subcarriers;     /* F = 12 * n_size_bwp: a positive multiple of focc_block, < 65536  */
    int32_t num_ofdm_symbols;    /* T: must be 14                                                   */
    int32_t d_s;                 /* state width: a multiple of 4 in [4, 60] (56 in all shipped configs; two
                                    more channels of the 64-wide state rows carry the positional encoding)  */
    int32_t num_it;              /* number of CGNNIt blocks in the weight file (>= 1)               */
    int32_t units_init[2];       /* num_units_init: two hidden layers, each in [1, 128]             */
    int32_t units_agg;           /* num_units_agg[i]: one hidden layer in [1, 64]                   */
    int32_t units_state[2];      /* num_units_state[i]: two hidden layers, each in [1, 128]; the
                                    same widths in every iteration                                  */
    int32_t units_readout;       /* num_units_readout: one hidden layer in [1, 128]                 */
    int32_t n_io;                /* number of StateInit / ReadoutLLRs stacks in the file, [1, NRX_MAX_IO] */
    int32_t io_bits[NRX_MAX_IO]; /* output width of each LLR head, [1, 16]                          */
    int32_t num_dmrs_symbols;
    int32_t dmrs_symbols[NRX_MAX_DMRS];
    int32_t focc_block;          /* 2 * num_cdm_groups_without_data                                 */
    int32_t num_data_res;        /* data REs per UE (RG demapper output length / bits)              */
} nrx_model_desc;

/* Replaces: building CGNNOFDM + `load_weights(model, path)` (utils/utils.py:53-70,
 * scripts/evaluate.py:185-187).  `weight_arrays` is the unpickled Keras get_weights() list in
 * file order (fp32, host memory), `weight_sizes[i]` its element counts; it is validated against
 * `desc` exactly like set_weights would.  Geometry tables (host memory):
 *   pilots     [U][n_dmrs*F][2]  float32  pilot_pattern.pilots (zeros off-comb)
 *   nn_index   [U][T*F]          int32    nearest-pilot gather (utils/neural_rx.py:973-992)
 *   pos_enc    [U][F][T][2]      float32  positional encoding (utils/onnx_utils.py:172-260)
 *   data_index [T*F]             int32    ordinal among data REs or -1 (RG demapper)
 * The engine copies everything to `device`; nothing is retained from the host pointers. */
int nrx_create(const nrx_model_desc* desc, const float* const* weight_arrays,
               const int64_t* weight_sizes, int32_t num_arrays, const float* pilots,
               const int32_t* nn_index, const float* pos_enc, const int32_t* data_index,
               int32_t device, nrx_engine** out);

int nrx_destroy(nrx_engine* e);

/* Replaces: the `num_it` property setter (utils/neural_rx.py:532-542): 1 <= n <= len(iterations),
 * otherwise NRX_ERR_INVALID "Invalid number of iterations". */
int nrx_set_num_it(nrx_engine* e, int32_t num_it);
int nrx_get_num_it(const nrx_engine* e, int32_t* num_it);

/* Slots pushed through all layers together (activations of one group stay L2-resident).
 * 0 = whole batch in one pass. */
int nrx_set_slots_per_pass(nrx_engine* e, int32_t slots);

/* Execution plan of the sep-conv stacks (StateInit :61-132, UpdateState :210-270):
 *   fused == 1:           one kernel per stack, the two 128-channel hidden activations stay in
 *                         shared memory (line-buffer fusion along the subcarrier axis); one
 *                         aggregation kernel per iteration (any number of users);
 *   fused == 2:           as 1, but with two users the message MLP of AggregateUserStates
 *                         (:184-188) runs in the tail of the preceding stack, each user reads the
 *                         other user's message tensor directly and no aggregation kernel is
 *                         launched (measured on B200: same speed as plan 1);
 *   fused == 5:           as 1, but the stacks run in the warp-specialised, software-pipelined kernel
 *                         (nrx_stack_ws.cuh: depthwise warps, an MMA / TMA issuing warp and epilogue
 *                         warps work on two tiles concurrently; needs cuTensorMapEncodeTiled from
 *                         the driver).  Bit-identical to plan 1; 3 % faster per 30-slot UpdateState
 *                         launch on B200, 7 % slower per StateInit launch;
 *   fused == 6 (default): StateInit in the serial kernel of plan 1, UpdateState in the pipelined kernel of
 *                         plan 5 — each stack in the kernel that runs it fastest (plan 1 if the driver lacks
 *                         cuTensorMapEncodeTiled);
 *   fused == 0:           one kernel per SeparableConv2D layer, activations through HBM/L2;
 *   fused == 3, 4:        two measured-slower experiments of round 1 (CTA-pair GEMMs, TMEM-resident
 *                         stack); only in builds with -DNRX_EXPERIMENTAL_PLANS, otherwise
 *                         NRX_ERR_UNSUPPORTED.
 * Plans 0, 1, 5 and 6 are bit-identical; plan 2 differs at fp16 round-off. */
int nrx_set_fused(nrx_engine* e, int32_t fused);

/* Inactive-user skipping (off by default = the reference's behaviour: it computes all users and ignores the
 * inactive ones, notebooks/nrx_architecture.ipynb:537).  An inactive user's state never reaches an active user
 * (its messages are masked, utils/neural_rx.py:192-204), so with skipping on the (slot, user) planes of inactive
 * users are not computed: the LLRs and the refined channel estimate of ACTIVE users are bit-identical, those of
 * inactive users are zeros instead of the reference's meaningless values.  With one of two users active the
 * work halves.  Plans 1 and 5 (ignored by the others). */
int nrx_set_skip_inactive(nrx_engine* e, int32_t enable);

/* Bytes of device scratch nrx_forward needs for `batch` slots. */
int nrx_workspace_bytes(const nrx_engine* e, int32_t batch, size_t* bytes);

/* Replaces: NeuralPUSCHReceiver.estimate_channel + CGNNOFDM.forward inference branch
 * (utils/neural_rx.py:1462-1514, 813-881; CGNN.forward :544-595).  All pointers are DEVICE
 * pointers owned by the caller; the call only enqueues work on `cuda_stream` (a cudaStream_t).
 *   io_index   [B][U] int32 or NULL: StateInit stack per UE (the one-hot mcs_ue_mask, :562-569);
 *              NULL = stack `llr_head` for every UE (mcs_ue_mask_eval=None, :818-823)
 *   head_index [B][U] int32 or NULL: LLR readout head per UE; NULL = head `llr_head` for every
 *              UE (the reference behaviour: only mcs_arr_eval[0] is evaluated, :847-858)
 *   out_bits   values written per RE, [1, 16]: io_bits[llr_head] normally, fewer in masking mode
 * io_index / head_index values outside [0, n_io) are clamped on the device (never an out-of-bounds read).
 *   llr / llr_grid / h_hat_refined / h_hat_ls: outputs, any may be NULL (not written).
 * The call neither allocates nor synchronises, so it can be recorded into a CUDA graph with
 * stream capture (batch-1 latency: 98 us replayed vs 124 us eager for nrx_rt on B200).        */
int nrx_forward(nrx_engine* e, void* cuda_stream, int32_t batch, const void* y,
                const float* active_tx, const int32_t* io_index, const int32_t* head_index,
                int32_t llr_head, int32_t out_bits, float* llr, float* llr_grid,
                float* h_hat_refined, float* h_hat_ls, void* workspace, size_t workspace_bytes);

/* Replaces: NeuralReceiverONNX.forward (utils/neural_rx.py:1773-1812) with NRPreprocessing
 * (:1614-1713) — the Aerial / TensorRT-shaped call whose seven inputs are the TRT bindings of
 * scripts/export_onnx.py:153-160.  nrx_set_aerial_dmrs takes the two integer inputs
 *   dmrs_ofdm_pos       [U][n_sym] int32 (host)  DMRS symbol indices per UE
 *   dmrs_subcarrier_pos [U][n_sc]  int32 (host)  non-zero pilot subcarriers inside one PRB per UE
 * and precomputes the per-PRB nearest-pilot table and positional encoding (they only change
 * with the DMRS configuration).  nrx_forward_aerial takes DEVICE pointers:
 *   rx_slot_real / rx_slot_imag  [B][F][T][N_rx] float32
 *   h_hat_real / h_hat_imag      [B][n_pilots][U][N_rx] float32  LS estimates at the non-zero
 *                                pilots, DMRS-symbol major (n_pilots = n_sym * F/12 * n_sc)
 *   active_dmrs_ports            [B][U] float32 0/1
 * outputs
 *   llr    [B][bits][U][F][T] float32 = MINUS the Sionna-convention LLR (:1809-1810)
 *   h_hat  [B][U][F][T][2*N_rx] float32 refined channel estimate.
 * Single-MCS models only (the reference: "no support for mixed MCS", :1796). */
int nrx_set_aerial_dmrs(nrx_engine* e, const int32_t* dmrs_ofdm_pos, int32_t n_sym,
                        const int32_t* dmrs_subcarrier_pos, int32_t n_sc);
int nrx_forward_aerial(nrx_engine* e, void* cuda_stream, int32_t batch, const float* rx_slot_real,
                       const float* rx_slot_imag, const float* h_hat_real, const float* h_hat_imag,
                       const float* active_dmrs_ports, float* llr, float* h_hat, void* workspace,
                       size_t workspace_bytes);

/* Same call with HOST buffers — the call a drop-in user of the reference's receiver makes with
 * NumPy arrays.  The batch flows through a 3-stage pipeline in chunks of nrx_set_host_chunk slots
 * (H2D of chunk i+1, kernels of chunk i and D2H of chunk i-1 overlap on three streams).  Pinned /
 * registered buffers are DMA'd in place; pageable ones are staged through engine-owned pinned
 * memory.  Returns after all requested outputs are in the caller's buffers. */
int nrx_forward_host(nrx_engine* e, int32_t batch, const void* y, const float* active_tx,
                     const int32_t* io_index, const int32_t* head_index, int32_t llr_head,
                     int32_t out_bits, float* llr, float* llr_grid, float* h_hat_refined,
                     float* h_hat_ls);

/* The same call, asynchronous: it only ENQUEUES the copies and kernels and returns a ticket; nrx_wait(ticket)
 * blocks until all outputs of that call are in the caller's buffers.  All buffers (inputs and outputs) must be
 * page-locked host memory (cudaHostAlloc / cudaHostRegister / pinned torch tensors) — NRX_ERR_INVALID otherwise —
 * and belong to the engine until nrx_wait returns.  Up to 4 calls may be in flight: the chunk pipeline runs on
 * across calls, so the copy-in of call n+1 overlaps the kernels and copy-outs of call n (a serving loop keeps two
 * calls in flight and never exposes a transfer).  Calls complete in order.  Replaces the same reference call as
 * nrx_forward_host; the reference itself is synchronous (utils/neural_rx.py:1544-1603). */
int nrx_forward_host_async(nrx_engine* e, int32_t batch, const void* y, const float* active_tx,
                           const int32_t* io_index, const int32_t* head_index, int32_t llr_head,
                           int32_t out_bits, float* llr, float* llr_grid, float* h_hat_refined,
                           float* h_hat_ls, int64_t* ticket);
int nrx_wait(nrx_engine* e, int64_t ticket);

/* Slots per pipeline chunk of the host-buffer calls (0 = default: nrx_forward_host a third of the batch, at most 16;
 * nrx_forward_host_async whole calls up to 32 slots). */
int nrx_set_host_chunk(nrx_engine* e, int32_t slots);

/* Number of kernels nrx_forward enqueues for `batch` slots with the current settings. */
int nrx_launches_per_forward(const nrx_engine* e, int32_t batch, int32_t* launches);

/* Host-only planning helper (pure function, no device needed; exported for tests and tooling): the stack kernels
 * cut every (slot, user) plane of `num_subcarriers` into that many chunks of consecutive subcarriers (one work
 * item each; minimises waves x steps on num_sms persistent CTAs). */
int nrx_plan_stack_chunks(int32_t planes, int32_t num_subcarriers, int32_t num_sms, int32_t* chunks_per_plane);

/* Host-only planning helper for the default ("balanced") work distribution of the fused stack kernels: the planes of
 * a launch are laid end to end (planes x num_subcarriers subcarriers) and CTA `cta` of `num_ctas` walks the
 * subcarriers [*first, *last) of that line, cut into one work item per plane it touches (every item recomputes a
 * 4-subcarrier run-in).  The ranges partition the line; CTAs beyond (planes x num_subcarriers) / 5 get none. */
int nrx_plan_stack_range(int32_t planes, int32_t num_subcarriers, int32_t num_ctas, int32_t cta, int64_t* first,
                         int64_t* last);

/* Algorithmic multiply-accumulates per user resource element for head `llr_head` at the current
 * num_it (sum of weight elements, biases excluded — SURVEY.md App. A.6). */
int nrx_mac_per_pixel(const nrx_engine* e, int32_t llr_head, int64_t* macs);

/* Per-kernel timing for the roofline report (bench.py): with profiling enabled nrx_forward brackets
 * every launch with CUDA events on the launching stream; nrx_get_profile synchronises, adds the
 * elapsed milliseconds and launch counts of each kernel class since the last call into
 * ms[NRX_NUM_KERNEL_CLASSES] / launches[NRX_NUM_KERNEL_CLASSES] and resets the record. */
#define NRX_K_POWER 0        /* input power partial sums                          */
#define NRX_K_PREP 1         /* LS + FOCC + NN interpolation + normalisation      */
#define NRX_K_SEP_IN 2       /* sep-conv  32 -> 128 (first StateInit layer)       */
#define NRX_K_SEP_HID 3      /* sep-conv 128 -> 128 (+ReLU)                       */
#define NRX_K_SEP_INIT_OUT 4 /* sep-conv 128 -> d_s, StateInit output             */
#define NRX_K_SEP_UPD_OUT 5  /* sep-conv 128 -> d_s, UpdateState output + residual */
#define NRX_K_AGG 6          /* AggregateUserStates MLP + cross-user reduction    */
#define NRX_K_READOUT 7      /* ReadoutLLRs + ReadoutChEst + RG demapping         */
#define NRX_K_STACK_INIT 8   /* fused StateInit stack (3 sep-convs, hidden tiles on chip)         */
#define NRX_K_STACK_UPD 9    /* fused UpdateState stack (3 sep-convs + residual)                 */
#define NRX_NUM_KERNEL_CLASSES 10
int nrx_set_profiling(nrx_engine* e, int32_t enable);
int nrx_get_profile(nrx_engine* e, double* ms, int64_t* launches);

/* Test hooks: run ONE kernel of the path on caller-provided DEVICE tensors in the engine's internal activation
 * layout (fp16, one 64-channel row per user resource element, row = ((slot*U + user)*F + f)*14 + t;
 * state rows = [s (d_s) | pos. encoding (2) | 0]).  They exist so that tests can drive the kernels with
 * random tensors at edge-case widths against the reference's AggregateUserStates / StateInit / UpdateState /
 * read-out blocks (utils/neural_rx.py:135-207, 61-132, 210-270, 309-404); the product path does not use them.
 *   nrx_debug_aggregate: a = AggregateUserStates_it(s, active_tx)
 *   nrx_debug_stack:     it >= 0: s_out = UpdateState_it(a, s) (residual included);
 *                        it <  0: s_out = StateInit_stack(z0), z0 rows of 32 channels [y | pe | h_ls | 0]
 *                        (execution plan 1, 5 or 6 as set by nrx_set_fused)
 *   nrx_debug_readout:   llr_grid [B][U][F][T][out_bits], h_hat_refined [B][U][F][T][2*N_rx] from s
 *   nrx_debug_option:    NRX_OPT_AGG_PIPELINED (default 1): 0 runs the one-tile-per-CTA aggregation kernel also for
 *                        two users (the pipelined kernel must reproduce it bit for bit);
 *                        NRX_OPT_STACK_BALANCED (default 1): 0 cuts every plane into nrx_plan_stack_chunks equal chunks
 *                        instead of the balanced CTA ranges of nrx_plan_stack_range (same outputs bit for bit) */
#define NRX_OPT_AGG_PIPELINED 1
#define NRX_OPT_STACK_BALANCED 2
int nrx_debug_option(nrx_engine* e, int32_t option, int32_t value);
int nrx_debug_aggregate(nrx_engine* e, void* cuda_stream, int32_t it, int32_t batch, const void* s_f16,
                        const float* active_tx, void* a_f16);
int nrx_debug_stack(nrx_engine* e, void* cuda_stream, int32_t it, int32_t stack, int32_t batch, const void* z0_f16,
                    const void* a_f16, const void* s_f16, void* s_out_f16);
int nrx_debug_readout(nrx_engine* e, void* cuda_stream, int32_t head, int32_t batch, int32_t out_bits,
                      const void* s_f16, float* llr_grid, float* h_hat_refined);

const char* nrx_last_error(void);
const char* nrx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NRX_B200_H */
