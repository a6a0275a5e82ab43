/*
 * b200fft -- C ABI of the B200-native FFT engine that drops in for the cuFFT binding used by
 * accelerate-fft's PTX backend.  Every entry point below replaces one symbol of Hackage
 * `cufft` (module Foreign.CUDA.FFT) at the call site cited (paths relative to
 * /root/reference/src/Data/Array/Accelerate/Math/FFT/LLVM/).
 *
 * Conventions (mirroring cuFFT so the Haskell shim is a one-line-per-call edit):
 *   - every call returns an int status, 0 == B200FFT_SUCCESS;
 *   - transforms are OUT-OF-PLACE (in != out), the input is never written;
 *   - data is interleaved complex (re,im) float or double, dense row-major, last extent
 *     contiguous (Type.hs:32, PTX.hs:99);
 *   - results are UN-NORMALISED in both directions (FFT.hs applies the Inverse scale);
 *   - plans are immutable after creation; b200fftExec is re-entrant, takes the stream as an
 *     argument (no setStream race, cf. PTX.hs:121) and only enqueues work -- it never
 *     synchronises the host;
 *   - the library uses the CUDA context current on the calling thread and never sets a device.
 * No cuFFT, no CPU fallback: if no GPU kernel exists for a request the call fails.
 */
#ifndef B200FFT_H
#define B200FFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200fft_plan_s* b200fftHandle;      /* replaces FFT.Handle  (PTX/Plans.hs:41,42,50,66) */
typedef void* b200fftStream;                        /* a CUstream / cudaStream_t */

/* transform type tags; values equal cuFFT's so `fromEnum t` hashing stays identical (PTX.hs:126-128) */
#define B200FFT_C2C 0x29
#define B200FFT_Z2Z 0x69
/* directions (PTX.hs:130-132 fftMode: Forward -> FORWARD, Reverse/Inverse -> INVERSE) */
#define B200FFT_FORWARD (-1)
#define B200FFT_INVERSE 1

enum {
  B200FFT_SUCCESS = 0,
  B200FFT_INVALID_PLAN = 1,
  B200FFT_ALLOC_FAILED = 2,
  B200FFT_INVALID_TYPE = 3,
  B200FFT_INVALID_VALUE = 4,
  B200FFT_INTERNAL_ERROR = 5,
  B200FFT_EXEC_FAILED = 6,
  B200FFT_INVALID_SIZE = 8,
  B200FFT_NO_DEVICE = 11,
  B200FFT_NOT_SUPPORTED = 16
};

/* FFT.plan1D n t batch            -- PTX.hs:141 */
int b200fftPlan1d(b200fftHandle* plan, int64_t n, int type, int64_t batch);
/* FFT.plan2D h w t                 -- PTX.hs:148 */
int b200fftPlan2d(b200fftHandle* plan, int64_t h, int64_t w, int type);
/* FFT.plan3D d h w t               -- PTX.hs:155 */
int b200fftPlan3d(b200fftHandle* plan, int64_t d, int64_t h, int64_t w, int type);
/* FFT.planMany [n] Nothing Nothing t batch (rank 1, contiguous, idist = odist = n) -- PTX.hs:162,169 */
int b200fftPlanMany1d(b200fftHandle* plan, int64_t n, int64_t batch, int type);

/* One axis of a dense array viewed as [outer][n][inner] (inner = element stride of the transformed axis).
 * Building block of the slab-decomposed multi-GPU 3D transform; no reference call site. */
int b200fftPlanAxis(b200fftHandle* plan, int64_t outer, int64_t n, int64_t inner, int type);

/* FFT.setStream + FFT.execC2C / FFT.execZ2Z  -- PTX.hs:119-124.
 * direction: B200FFT_FORWARD or B200FFT_INVERSE; the element type is baked into the plan. */
int b200fftExec(b200fftHandle plan, const void* in, void* out, int direction, b200fftStream stream);
/* Same, with every output element multiplied by `scale` in the last butterfly pass
 * (the fused-Inverse entry of SURVEY.md section 8f-2; no reference call site yet). */
int b200fftExecScaled(b200fftHandle plan, const void* in, void* out, int direction, double scale,
                      b200fftStream stream);

/* The transform followed by DFT/Centre.hs's shift1D/2D/3D along every transformed axis, the half rotation folded into
 * the stores of each axis' last butterfly pass (no extra pass).  Even extents: also
 * fft(centre(x)) (Centre.hs:17-19).  B200FFT_NOT_SUPPORTED for extents that are not powers of two. */
int b200fftExecShifted(b200fftHandle plan, const void* in, void* out, int direction, double scale,
                       b200fftStream stream);

/* FFT.destroy -- PTX/Plans.hs:80.  Safe from any thread (GC finaliser). */
int b200fftDestroy(b200fftHandle plan);

const char* b200fftErrorString(int status);

/* Per-exec scratch of multi-pass plans is stream-ordered and comes from a library-owned memory pool that
 * keeps its pages between execs; this returns them to the driver (no reference counterpart). */
int b200fftTrimScratch(void);

/* Introspection used by the bench / tests (no reference counterpart). */
size_t b200fftScratchBytes(b200fftHandle plan);      /* stream-ordered scratch one exec allocates */
int b200fftNumPasses(b200fftHandle plan);            /* kernel launches (= HBM passes) per exec */
int64_t b200fftKernelLaunches(void);                 /* process-wide count of kernels launched */
int b200fftHasExperimental(void);                    /* 1 if built with B200FFT_EXPERIMENTAL=1 (opt-in kernels that measured slower) */
/* fills `buf` with a one-line-per-pass description of the plan; returns bytes written */
int b200fftDescribe(b200fftHandle plan, char* buf, int buflen);

/* Slab-decomposed 3D FFT across the GPUs of one box (new capability, the reference is
 * single-device).  Rank g owns z-planes [g*D/P,(g+1)*D/P).  Pack re-orders a local [dl][h][w] slab into P
 * contiguous peer blocks [P][dl][h/P][w] for the all-to-all; Unpack is its inverse. */
int b200fftSlabPack(int type, const void* src, void* dst, int64_t dl, int64_t h, int64_t w, int nranks, b200fftStream stream);
int b200fftSlabUnpack(int type, const void* src, void* dst, int64_t dl, int64_t h, int64_t w, int nranks, b200fftStream stream);

/* The exchange folded into the pass: one strided-axis plan (b200fftPlanAxis with n a power of two <= 2048) whose
 * stores are scattered over `npeers` buffers -- output index k of the transformed axis goes to outs[k / (n/npeers)]
 * at local index k % (n/npeers); inside a buffer the element of (outer o, local index kl, inner i) sits at
 * o*out_outer_stride + kl*out_n_stride + i (strides in complex elements).  The buffers may live on other GPUs of
 * the box (peer-mapped over NVLink): the y-axis pass of the slab-decomposed fft3D writes every ky row straight into
 * the memory of the rank that owns it, so pack + all-to-all + unpack disappear.  The caller orders the exchange
 * (a barrier on all ranks before the buffers are overwritten and after the kernel has finished). */
int b200fftExecScatter(b200fftHandle plan, const void* in, void* const* outs, int npeers, int64_t out_outer_stride,
                       int64_t out_n_stride, int direction, double scale, b200fftStream stream);
/* Same; max_ctas > 0 runs the pass as a grid-stride loop of at most that many CTAs, so that an NVLink-bound scatter pass
 * occupies only that many SMs and HBM-bound passes on other streams run beside it (0 = one CTA per tile). */
int b200fftExecScatterOn(b200fftHandle plan, const void* in, void* const* outs, int npeers, int64_t out_outer_stride,
                         int64_t out_n_stride, int direction, double scale, int max_ctas, b200fftStream stream);
/* b200fftPlanAxis through a window: element (o, k, i) of the axis sits at o*outer_stride + k*n_stride + i with i < inner,
 * i.e. `inner` may be a sub-range of the real rows (a chunk of columns).  Power-of-two n <= 2048 (one in-place pass). */
int b200fftPlanAxisView(b200fftHandle* plan, int64_t outer, int64_t n, int64_t inner, int64_t outer_stride, int64_t n_stride,
                        int type);
/* Device buffers another process of the same box can map (CUDA IPC, 64-byte handles exchanged by the caller). */
int b200fftPeerAlloc(void** ptr, size_t bytes);
int b200fftPeerFree(void* ptr);
int b200fftPeerExport(void* ptr, unsigned char handle[64]);
int b200fftPeerOpen(const unsigned char handle[64], void** ptr);
int b200fftPeerClose(void* ptr);

/*
 * Multi-GPU entry point: the slab-decomposed 3D transform across the GPUs of one NVLink box (SURVEY.md section 8b
 * "multi-GPU entry points ... additional exports", 8e; BASELINE config 5).  One process per GPU, the calling thread has
 * its device current.  Rank g owns z-planes [g*D/P, (g+1)*D/P) of the dense (D, H, W) array: `in` is that [D/P][H][W] slab.
 * The exchange is folded into the y pass's stores (peer memory over NVLink, CUDA IPC), ranks meet in flag barriers in peer
 * memory; everything is enqueued on `stream` and on library-owned side streams that fork from and join it -- no host
 * synchronisation, no collective library on the data path (csrc/slab.cu).  Power-of-two D, H <= 2048, divisible by nranks.
 *
 * Plan creation is collective.  The only thing the library cannot do itself is tell the ranks about each other: the host
 * supplies an all-gather of small byte blobs (MPI_Allgather, ncclAllGather + copies, torch.distributed.all_gather, a pipe),
 * called twice.  It returns 0 on success; `recv` holds nranks * bytes, rank r's blob at r * bytes.
 */
typedef struct b200fft_slab_s* b200fftSlabHandle;
typedef int (*b200fftAllgatherFn)(void* ctx, const void* send, void* recv, size_t bytes);
enum { B200FFT_SLAB_NATURAL = 1 };                 /* plan flag: also allocate what natural-layout output needs */
enum { B200FFT_SLAB_NATURAL_OUT = 0,               /* out = [D/P][H][W], this rank's z-slab of fft3D's result (FFT.hs:150-173) */
       B200FFT_SLAB_TRANSPOSED_OUT = 1 };          /* out = [H/P][D][W]: this rank's ky rows, each with all kz (one exchange less) */
int b200fftPlanSlab3d(b200fftSlabHandle* plan, int64_t d, int64_t h, int64_t w, int type, int rank, int nranks, int flags,
                      b200fftAllgatherFn allgather, void* ctx);
/* Collective: every rank calls it with the same arguments in the same order.  `scale` multiplies the result in the last
 * pass (1/(D*H*W) for Mode Inverse, FFT.hs:155,172).  `in` is not written; `out` must not alias it.  Unlike b200fftExec, a
 * slab plan owns per-transform state (receive buffers, barrier epochs): one transform at a time per plan -- calls on one
 * plan must come from one host thread at a time and are ordered on the device by the library's own barriers.  An error on
 * one rank (plan creation or exec) leaves the others waiting in the all-gather / at a barrier, as with any collective. */
int b200fftExecSlab(b200fftSlabHandle plan, const void* in_slab, void* out, int direction, double scale, int layout,
                    b200fftStream stream);
/* Natural layout: the result is assembled by the peers in a library-owned buffer and then copied to `out`; passing THIS
 * buffer as `out` skips the copy (valid until the next exec on any rank has started). */
int b200fftSlabNaturalBuffer(b200fftSlabHandle plan, void** ptr);
/* Pipelining knobs of one output layout (synchronises the device): the y pass goes in col_chunks column chunks x
 * plane_chunks plane chunks on y_ctas CTAs (0 = one per tile); z of a chunk runs beside y of the next, x of a plane chunk
 * beside y of the previous.  The defaults are the measured optimum on 8 x B200. */
int b200fftSlabTune(b200fftSlabHandle plan, int layout, int plane_chunks, int col_chunks, int y_ctas);
/* Collective by contract: no rank may destroy while another still executes. */
int b200fftDestroySlab(b200fftSlabHandle plan);

/*
 * Host-side mirror of the reference's public API for this path (FFT.hs:63-173 + PTX.hs +
 * PTX/Plans.hs): shape/rank dispatch, Mode semantics, the five plan caches and the Inverse
 * normalisation.  `mode`: 0 Forward, 1 Reverse, 2 Inverse (Mode.hs:15-19).
 * `rank`/`shape`: Accelerate shape, outermost first, shape[rank-1] contiguous.
 * Device-pointer flavour (what foreignAcc would call) ...
 */
int accfft_fft(int mode, int rank, const int64_t* shape, int type, const void* d_in, void* d_out,
               b200fftStream stream);                                   /* FFT.hs:63-84  / PTX.hs:52-61 */
int accfft_fft1D(int mode, int64_t n, int type, const void* d_in, void* d_out, b200fftStream stream);   /* FFT.hs:92-111  */
int accfft_fft2D(int mode, int64_t h, int64_t w, int type, const void* d_in, void* d_out, b200fftStream stream); /* FFT.hs:119-142 */
int accfft_fft3D(int mode, int64_t d, int64_t h, int64_t w, int type, const void* d_in, void* d_out,
                 b200fftStream stream);                                 /* FFT.hs:150-173 */
/* ... and host-buffer flavour (what `run` does around it): H2D copy, transform, D2H copy,
 * synchronous.  kind: 0 fft (innermost axis), 1 fft1D, 2 fft2D, 3 fft3D. */
int accfft_run_host(int kind, int mode, int rank, const int64_t* shape, int type, const void* h_in, void* h_out);
/* Same for a chain of `nmodes` transforms applied one after the other with the array staying on the device in
 * between (e.g. {0, 2} = Inverse . Forward) -- one Accelerate `run` holding several Aforeign nodes.  For kind 0
 * the rows are independent, so the array flows through the device in chunks of whole rows on three streams:
 * H2D of chunk c+1, the kernels of chunk c and D2H of chunk c-1 overlap.  Pass pinned host buffers. */
int accfft_run_host_seq(int kind, const int* modes, int nmodes, int rank, const int64_t* shape, int type,
                        const void* h_in, void* h_out);
/* DFT/Centre.hs (rank 1..3): centre1D/2D/3D (:36-66), shift1D..3D and ishift1D..3D (:70-164; inverse != 0 selects the ishift family) as
 * stand-alone passes, and shiftND(fftND mode x) in one call (kind 1/2/3) -- fused into the transform's last passes
 * for power-of-two extents, transform + stand-alone shift otherwise. */
int accfft_centre(int rank, const int64_t* shape, int type, const void* d_in, void* d_out, b200fftStream stream);
int accfft_shift(int rank, const int64_t* shape, int type, int inverse, const void* d_in, void* d_out, b200fftStream stream);
int accfft_fft_centred(int kind, int mode, const int64_t* shape, int type, const void* d_in, void* d_out, b200fftStream stream);
/* when set non-zero the Inverse scale is fused into the last pass instead of a second kernel */
void accfft_set_fused_inverse(int on);
int accfft_plan_cache_size(void);
void accfft_plan_cache_clear(void);

#ifdef __cplusplus
}
#endif
#endif /* B200FFT_H */
