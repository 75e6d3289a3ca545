/*
 * swn_b200.h -- C ABI of the B200-native shallow-wavenet generation path.
 *
 * The reference (patrickltobing/shallow-wavenet) has no FFI layer: its boundary for this
 * path is the Python class API of src/nets.  This header is the boundary a
 * non-Python host binds instead; the Python mirror of the reference API
 * (shallow-wavenet_b200/nets/{cswnv_shift1,dswnv}.py) is a thin ctypes caller of exactly
 * these entry points.  Each entry point cites the reference interface it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; swn_last_error() gives the
 *     message of the last failing call on the calling thread;
 *   - a handle is bound to one CUDA device, is not thread-safe, and owns every device
 *     allocation it makes;
 *   - "dev" pointers are device pointers owned by the caller (e.g. torch tensors), "host"
 *     pointers are host memory; `stream` is a cudaStream_t passed as void* (NULL = default);
 *   - there is NO CPU fallback: on a device that is not sm_100 (B200) swn_create fails.
 */
#ifndef SWN_B200_H
#define SWN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWN_ABI_VERSION 5

typedef struct swn_model* swn_handle;

enum { SWN_KIND_CSWNV = 0,      /* Laplace / LPC vocoder, src/nets/cswnv_shift1.py:130 */
       SWN_KIND_DSWNV = 1 };    /* mu-law softmax vocoder, src/nets/dswnv.py:190       */

/* Constructor arguments: CSWNV.__init__ (src/nets/cswnv_shift1.py:131-133) and
 * DSWNV.__init__ (src/nets/dswnv.py:191-193).  Unused fields (seg/lpc/aux_conv2d_flag for
 * DSWNV; n_quantize/audio_in_flag for CSWNV) are ignored. */
typedef struct swn_config {
    int32_t kind;
    int32_t n_quantize;
    int32_t n_aux;
    int32_t hid_chn;
    int32_t skip_chn;
    int32_t aux_kernel_size;
    int32_t aux_dilation_size;
    int32_t dilation_depth;
    int32_t dilation_repeat;
    int32_t kernel_size;
    int32_t upsampling_factor;
    int32_t seg;
    int32_t lpc;
    int32_t aux_conv2d_flag;
    int32_t wav_conv_flag;
    int32_t audio_in_flag;
} swn_config;

/* What swn_generate does.  One call = conditioning already done (swn_condition), then the
 * warm-up over the padded seed and the whole sampling loop of
 * CSWNV.batch_fast_generate (src/nets/cswnv_shift1.py:314-402) /
 * DSWNV.batch_fast_generate (src/nets/dswnv.py:312-374) as ONE persistent kernel. */
typedef struct swn_gen_args {
    int64_t n_samples;        /* max(n_samples_list); generated = (n_samples/seg)*seg per utterance */
    uint64_t seed;            /* Philox4x32-10 key when noise_dev == NULL */
    const float* noise_dev;   /* optional injected noise (parity tests):
                                 CSWNV eps [B][n_gen] in (-0.4999,0.5)   (cswnv_shift1.py:373)
                                 DSWNV q   [B][n_gen][Q] ~ Exp(1)        (multinomial fast path, dswnv.py:365) */
    const void* forced_dev;   /* optional teacher forcing: samples fed back instead of the sampled
                                 ones; float [B][n_gen] (CSWNV) / int64 [B][n_gen] (DSWNV) */
    void* out_dev;            /* generated samples, float [B][n_gen] / int64 [B][n_gen] */
    float* head_dev;          /* optional raw out_2 rows per step, [B][n_gen/seg][head_dim]
                                 (head_dim = 2*seg+lpc or n_quantize) */
    const void* seed_dev;     /* optional seed waveform (the `audio` argument of batch_fast_generate):
                                 float [B][seg] (CSWNV) / int64 [B][1] (DSWNV); NULL = the decode
                                 scripts' seed (zeros / n_quantize/2, decode_cswnv_laplace-shift1.py:93,
                                 decode_dswnv_softmax.py:94-99) */
    int32_t batch_tile;       /* utterances processed together per pipeline slot; 0 = auto */
    int32_t poll_mode;        /* tuning: 0 = default, 1 = pure per-word polling, 2 = sentinel + per-word */
    uint64_t* prof_dev;       /* optional instrumentation: [prof_count][grid][32] %globaltimer stamps */
    int32_t prof_start;       /* first loop iteration stamped */
    int32_t prof_count;
    int32_t kernel_sel;       /* 0 = auto (small batches: cluster / layer-group kernel, else the pipelined batch kernel; table in INTEGRATION.md),
                                 1 = force the layer-group kernel, 2 = force the cluster kernel,
                                 3 = force the batch kernel (tensor-core cells where they fit), 4 = batch kernel, FFMA cells */
    int32_t debug_helper_delay_ns; /* test hook, 0 in production: every helper CTA of the cluster kernels sleeps this long per
                                 turn, so that the core has to wait for the helper sums (tests/test_gpu_parity.py holds the
                                 kernel to bit-identical output under such delays) */
    const int64_t* n_samples_list; /* optional HOST array [B]: samples wanted per utterance (the reference's
                                 n_samples_list, cswnv_shift1.py:291-295,426-428).  The batch kernel stops a pipeline
                                 slot once all its utterances are complete, the seg cluster kernel runs the (up to three)
                                 utterances that share a launch to the longest of them; rows of out_dev are only defined up to each
                                 utterance's own length (rounded up to seg).  NULL = n_samples for everyone. */
} swn_gen_args;

typedef struct swn_info {
    int32_t abi_version;
    int32_t receptive_field;  /* cswnv_shift1.py:182 */
    int32_t eff_seg;          /* samples per loop step */
    int32_t head_dim;
    int32_t n_sm;             /* SMs used by the persistent kernel */
    int32_t ctas_per_layer;   /* variant 1, 3, 4: CTAs per layer group; variant 2: CTAs of the core cluster (16) */
    int32_t head_ctas;        /* variant 1: head CTAs; variant 2: helper CTAs; variant 3, 4: dense head stage + tail CTAs */
    int32_t smem_bytes;       /* dynamic shared memory per CTA of the last launch */
    int32_t batch_tile;       /* of the last launch */
    int32_t n_slots;          /* pipeline slots (ceil(B / batch_tile)) of the last launch */
    int64_t weight_bytes;     /* packed in-loop weights resident on the device */
    int64_t cond_bytes;       /* conditioning table bytes of the last swn_condition */
    float last_cond_ms;       /* device time of the last swn_condition (CUDA events) */
    float last_gen_ms;        /* device time of the last swn_generate persistent kernel */
    int64_t kernel_launches;  /* kernels launched by this handle so far */
    int32_t grid;             /* CTAs of the last persistent launch */
    int32_t kernel_variant;   /* 1 = layer groups over the chip, 2 = core cluster + helpers,
                                 3 = pipelined batch kernel with tcgen05 (tf32 x3) cells, 4 = the same with fp32 FFMA cells,
                                 5 = swn_forward (time-parallel tcgen05 GEMMs; last_gen_ms is then the forward's device time) */
} swn_info;

/* CSWNV(...) / DSWNV(...) followed by model.cuda() (decode_cswnv_laplace-shift1.py:208-223,
 * decode_dswnv_softmax.py:216-230). */
int swn_create(const swn_config* cfg, int device, swn_handle* out);

/* model.load_state_dict(...) (decode_cswnv_laplace-shift1.py:224): one call per state_dict
 * entry, `key` is the reference parameter name, `host` is contiguous
 * fp32 in the reference's own shape.  Shapes are checked against the constructor arguments. */
int swn_set_tensor(swn_handle h, const char* key, const float* host, const int64_t* shape, int32_t ndim);

/* Checks that every parameter was supplied, folds/packs the weights into the kernel layout
 * and uploads them.  Must be called after the last swn_set_tensor and before swn_condition. */
int swn_finalize(swn_handle h);

/* x = upsampling(conv_aux(scale_in(aux))) + replicate pad + seg window + in_x[l] for all
 * layers (cswnv_shift1.py:297-311,282; dswnv.py:302-309,291), evaluated once per batch at
 * frame rate.  aux_dev: fp32 [B][n_aux][T_f] (zero padded to the longest utterance exactly
 * as decode_cswnv_laplace-shift1.py:30-48 does). */
int swn_condition(swn_handle h, const float* aux_dev, int32_t B, int32_t T_f, void* stream);

/* The hot loop.  Requires a preceding swn_condition with the same B.  Asynchronous on `stream`
 * (argument validation aside) in the steady state of a decode loop, i.e. while the PLAN stays
 * the same.  The plan - CTA roles, packed per-CTA weights, exchange buffers - depends on
 * (kernel variant, B, batch_tile, T_f class); the first call and every call that changes it
 * synchronises `stream`, (re)allocates device scratch with cudaMalloc and uploads descriptors
 * before launching.  The batch kernel (variants 3, 4) additionally synchronises `stream` on
 * every call that passes n_samples_list, because the per-slot lengths live in one host-written
 * device buffer.  Nothing is allocated or synchronised when a decode loop repeats the same
 * (B, T_f, n_samples) shape, which is what bench.py times. */
int swn_generate(swn_handle h, const swn_gen_args* args, void* stream);

/* Teacher-forced evaluation of the same recurrence (SURVEY.md 8b `swn_teacher_forced`): the given samples are fed
 * back instead of sampled ones and the raw out_2 rows of every loop step are returned - the per-step outputs that
 * the parity tests compare with the reference module run on the same samples (a forward hook on out_2 around
 * cswnv_shift1.py:367 / dswnv.py:361).  Time alignment is that of batch_fast_generate (warm-up over the padded
 * seed, cswnv_shift1.py:314-334), not of the training-time forward() (cswnv_shift1.py:191-267).
 * samples_dev: float [B][n_gen] (CSWNV) / int64 [B][n_gen] (DSWNV); head_dev: float [B][n_gen/seg][head_dim];
 * n_gen = (n_samples/seg)*seg.  Requires a preceding swn_condition with the same B.  Asynchronous on `stream`;
 * equivalent to swn_generate with forced_dev / head_dev set and the sampled output discarded. */
int swn_teacher_forced(swn_handle h, const void* samples_dev, int64_t n_samples, float* head_dev, void* stream);

/* Arguments of swn_forward. */
typedef struct swn_fwd_args {
    const void* audio_dev;    /* CSWNV: float [B][T_a] waveform; DSWNV: int64 class indices [B][T_a], or, when
                                 audio_onehot != 0, the reference's float one-hot tensor [B][n_quantize][T_a] */
    int32_t audio_onehot;
    int32_t clip;             /* CSWNV forward(..., clip=True): also fill b_clip_dev / log_b_clip_dev */
    int64_t T_a;              /* must be T_f * upsampling_factor - seg (seg = 1 for DSWNV) */
    float* out_dev;           /* raw out_2 rows, [B][T_out][head_dim], T_out = T_f * U - 2 * seg + 1 */
    float* b_dev;             /* CSWNV: exp(logsigmoid(out[seg:2seg])) [B][T_out][seg] */
    float* log_b_dev;         /* CSWNV: logsigmoid(out[seg:2seg])      [B][T_out][seg] */
    float* b_clip_dev;        /* CSWNV, clip: exp(clamp(log_b, log 7.07e-7)) */
    float* log_b_clip_dev;
} swn_fwd_args;

/* Training-time / evaluation forward of the modules: CSWNV.forward(aux, audio)
 * (src/nets/cswnv_shift1.py:191-267) and DSWNV.forward(audio, aux) (src/nets/dswnv.py:250-276) - every time step of
 * whole utterances at once, zero conv history before the first sample (no warm-up), stride-1 outputs.  Requires a
 * preceding swn_condition(aux) with the same batch.  Time-parallel: the dilated convs and the 1x1 projections run as
 * tcgen05 (tf32 x 3) GEMMs over all time steps.  Inference only (no gradients; dropout `do=True` is not available).
 * Asynchronous on `stream` once its buffers exist (they grow with B * T_out: ~20 KB per time step). */
int swn_forward(swn_handle h, const swn_fwd_args* args, void* stream);

/* Blocks until the stream is idle and reports a device-side failure of the last
 * swn_generate (e.g. the in-kernel watchdog fired) as a non-zero status. */
int swn_sync(swn_handle h, void* stream);

/* End-to-end convenience for hosts without a device allocator: host aux in, host samples
 * out (H2D copy, swn_condition, swn_generate, D2H copy, synchronise).
 * out_host: float [B][n_gen] (CSWNV) or int64 [B][n_gen] (DSWNV). */
int swn_generate_host(swn_handle h, const float* aux_host, int32_t B, int32_t T_f, int64_t n_samples,
                      uint64_t seed, void* out_host);

int swn_get_info(swn_handle h, swn_info* info);

/* Philox4x32-10 stream used by the sampler, for host-side reproduction in tests:
 * fills out[n] with the uniforms the kernel draws for utterance b, samples n0..n0+n-1
 * (CSWNV: one per sample; DSWNV: Q per sample, row-major). */
int swn_philox_uniform(uint64_t seed, int32_t kind, int32_t n_quantize, int64_t b, int64_t n0, int64_t n, float* out_host);

/* Measurement aid (no reference counterpart): aggregate shared-memory read bandwidth of the
 * device in GB/s, the denominator of the batch-1 step roofline. */
int swn_measure_smem_bandwidth(int device, double* gb_per_s);

/* Noise-shaping restore filter over a decoded batch (recipe stages 6 / 9): what
 * src/bin/noise_shaping.py:52-87 (synthesis_diff: pysptk mc2b + MLSADF(order, alpha, pd) + Synthesizer with the frame
 * shift `hop`) followed by src/bin/feature_extract.py:57-76 (low_cut_filter: lfilter(firwin(255, ...), 1, x)) compute,
 * in float64 as the reference does.  Not tied to a model handle.  The trailing partial frame is not synthesised (pysptk
 * leaves it uninitialised; zeros here).  Synchronous with respect to its host arguments; the kernels run on `stream`. */
typedef struct swn_ns_args {
    const double* x_dev;          /* [B][stride] waveforms                                                       */
    double* y_dev;                /* [B][stride] filtered waveforms (may not alias x_dev)                        */
    double* tmp_dev;              /* [B][stride] scratch between the two filters (unused when fir_taps == 0)     */
    int64_t stride;
    int32_t B;
    int32_t order;                /* mel-cepstrum order m = dim - 1 (2 .. 62)                                    */
    const int64_t* n_samples;     /* host [B]: valid samples of every waveform                                   */
    const int32_t* n_frames;      /* host [B]: coefficient frames of every waveform (pyworld.harvest's count,
                                     noise_shaping.py:160-161); NULL = mcep_frames                               */
    const double* mcep_host;      /* [mcep_utts][mcep_frames][order + 1] differential mel-cepstrum (:165-170);
                                     NULL = no MLSA stage, the low-cut FIR alone (x_dev -> y_dev)                */
    int32_t mcep_utts;            /* 1 = shared by all waveforms, else B                                         */
    int32_t mcep_frames;          /* 1 = one vector for every frame (the script tiles the mean vector, :170)     */
    int32_t pade;                 /* 4 (pysptk's default) or 5                                                   */
    int32_t hop;                  /* int(fs / 1000 * shiftms) (:76)                                              */
    double alpha;                 /* all-pass constant (--mcep_alpha)                                            */
    const double* fir_host;       /* low-cut FIR taps (feature_extract.py:73) or NULL                            */
    int32_t fir_taps;
    int32_t reserved;
} swn_ns_args;
int swn_noise_shaping(int device, const swn_ns_args* args, void* stream);

int swn_destroy(swn_handle h);

const char* swn_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SWN_B200_H */
