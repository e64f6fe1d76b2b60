/* ias_b200.h -- C ABI of libias_b200.so: the B200 (sm_100a) front end of inverse-audio-synthesis.
 *
 * The reference (turian/inverse-audio-synthesis) has no FFI of its own: its hot path is three Python class
 * surfaces.  Each entry point below is the device-side replacement a binding for one of those
 * surfaces calls; the reference interface it replaces is cited as file:line into /root/reference (torchsynth call
 * sites where the code lives in the un-vendored torchsynth package).  ias_b200/{voice,pqmf,vicreg}.py are the
 * ctypes bindings; INTEGRATION.md shows the stub a maintainer adds to the reference.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *   - tensors are contiguous row-major fp32; pointers 16-byte aligned;
 *   - nothing is allocated per call: scratch comes in through (workspace, workspace_bytes);
 *   - work is enqueued on `stream` (a cudaStream_t) with no hidden synchronisation;
 *   - return 0 on success, an IAS_ERR_* code otherwise, message in ias_last_error() (thread local);
 *   - never throws, never exits.  One host thread per device at a time.
 */
#ifndef IAS_B200_H
#define IAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IAS_OK 0
#define IAS_ERR_INVALID 1     /* bad argument */
#define IAS_ERR_CUDA 2        /* CUDA runtime error (launch, memcpy) */
#define IAS_ERR_UNSUPPORTED 3 /* shape outside what the kernels cover */
#define IAS_ERR_WORKSPACE 4   /* workspace missing or too small */
#define IAS_ERR_NCCL 5        /* NCCL error */

#define IAS_VOICE_NPARAMS 78   /* conf/config.yaml:27 */
#define IAS_VOICE_NCONTROL 5   /* vco_1_pitch, vco_1_amp, vco_2_pitch, vco_2_amp, noise_amp */
#define IAS_NCCL_ID_BYTES 128

typedef void* ias_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define IAS_API __attribute__((visibility("default")))
#else
#define IAS_API
#endif

/* ---- library ---------------------------------------------------------------------------------------- */
IAS_API int ias_version(void);
IAS_API const char* ias_last_error(void);
/* IAS_OK iff `device` is a compute-capability 10.x GPU (the only target). */
IAS_API int ias_device_check(int device);

/* Launch counter and optional per-kernel CUDA-event timing (bench.py's roofline leg).  Every kernel launch the
 * library makes is counted; with profiling enabled each launch is also bracketed by an event pair on its stream. */
IAS_API int ias_prof_enable(int on);
IAS_API int ias_prof_reset(void);
IAS_API int ias_prof_kernel_count(void);
IAS_API const char* ias_prof_kernel_name(int id);
IAS_API long long ias_prof_launches(int id); /* id < 0: all kernels */
IAS_API int ias_prof_read(int id, double* total_ms, long long* timed_launches); /* synchronises on the events */

/* ---- Voice: torchsynth.synth.Voice (vicreg_audio_params.py:86-94,114; audio_to_params.py:196-203,215,238-257) --
 * Parameter block layout used by every voice entry point: params01[78][B], parameter-major, rows in torchsynth
 * *registration* order (= nn.Module.parameters() order = columns of the `params` tensor Voice.forward returns). */

/* "module/param" of registration row `reg_index` (0..77), NULL if out of range. */
IAS_API const char* ias_voice_param_name(int reg_index);
/* Position of registration row `reg_index` in sorted(named_parameters()), the order randomize(seed) assigns. */
IAS_API int ias_voice_sorted_index(int reg_index);

/* AbstractSynth.randomize(seed) + _batch_idx_to_is_train: sound i of the batch gets the first 78 outputs of
 * torch's CPU MT19937 seeded with (first_sound_id + i) (bit-identical to torch.rand(78, generator=g)).
 * Rows whose frozen78_host[row] != 0 are left untouched (freeze_parameters); frozen78_host may be NULL.
 * is_train[B] (uint8) may be NULL.  first_sound_id = batch_idx * batch_size. */
IAS_API int ias_voice_seed_params(int64_t first_sound_id, int B, const uint8_t* frozen78_host, float* params01,
                          uint8_t* is_train, ias_stream_t stream);

/* Same, with the batch number read from device memory when the kernel runs (first_sound_id = *batch_idx_dev * B):
 * the step can be captured in a CUDA graph and replayed with a new batch number, and the host never has to read the
 * DataLoader's integer back (the reference's `batch.cpu()` at vicreg_audio_params.py:109-112 is a synchronisation). */
IAS_API int ias_voice_seed_params_dev(const int64_t* batch_idx_dev, int B, const uint8_t* frozen78_host, float* params01,
                              uint8_t* is_train, ias_stream_t stream);

IAS_API size_t ias_voice_workspace_bytes(int B, int T, int C);

/* Control-rate stage only (keyboard, 6 ADSR, 2 LFO, modulation matrix): ctrl[B][5][C].  Inspection entry point
 * used by the parity tests; ias_voice_render runs the same kernel internally. */
IAS_API int ias_voice_control(const float* params01, int B, int C, float control_rate, float eps, float* ctrl,
                      void* workspace, size_t workspace_bytes, ias_stream_t stream);

/* Voice.output(): params -> audio[B][T].  noise[noise_rows][T] is the Noise module buffer (row b uses
 * noise[b % noise_rows]).  peak[B] receives max|mixed| before normalisation (may be NULL).  normalize == 1 applies
 * util.normalize_if_clipping (x / peak where peak > 1) -- a second pass over the clipping rows; normalize == 0 leaves
 * the raw mix; normalize == 2 defers the normalisation to the consumer: audio is the raw mix and peak[B] receives the
 * factor to apply instead (RN(1/peak) for a clipping row, else 1) -- exactly the row_scale argument of
 * ias_pqmf_analysis*, so synth -> PQMF needs no second pass over the audio.  Two inspection hooks for the parity tests, both normally NULL: ctrl_in[B][5][C] replaces the
 * control-rate signals the audio stage reads (the per-voice constants still come from params01), and phase_dbg
 * receives the two VCO cosine arguments, [B][2][T]. */
IAS_API int ias_voice_render(const float* params01, const float* noise, int noise_rows, float* audio, float* peak, int B,
                     int T, int C, float sample_rate, float control_rate, float eps, int normalize,
                     const float* ctrl_in, float* phase_dbg, void* workspace, size_t workspace_bytes,
                     ias_stream_t stream);

/* The same call split in two, so that a caller can overlap the (latency-bound) control stage of the NEXT batch with
 * whatever consumes the current audio: IAS_VOICE_STAGE_CONTROL runs ADSR + LFO/modulation + the work-queue schedule
 * from params01 into `workspace` (noise / audio / peak may be NULL); IAS_VOICE_STAGE_AUDIO renders the audio from a
 * workspace a control-stage call filled (params01 / ctrl_in unused).  Both bits = ias_voice_render.
 * The control stage itself splits once more, for callers that pipeline two batches deep: IAS_VOICE_STAGE_ENVELOPES
 * runs the six ADSR envelopes alone (compute bound: SLEEF-exact pow per control point), IAS_VOICE_STAGE_MODULATION
 * the LFOs / modulation matrix / per-interval records / schedule from the envelopes a previous call left in the same
 * workspace and the same params01.  ENVELOPES | MODULATION = CONTROL. */
#define IAS_VOICE_STAGE_CONTROL 1
#define IAS_VOICE_STAGE_AUDIO 2
#define IAS_VOICE_STAGE_ENVELOPES 4
#define IAS_VOICE_STAGE_MODULATION 8
IAS_API int ias_voice_render_stages(const float* params01, const float* noise, int noise_rows, float* audio, float* peak,
                            int B, int T, int C, float sample_rate, float control_rate, float eps, int normalize,
                            const float* ctrl_in, float* phase_dbg, void* workspace, size_t workspace_bytes, int stages,
                            ias_stream_t stream);

/* ---- PQMF: pqmf.PQMF (pqmf.py:9-55; callers audioembed.py:38, vicreg_audio_params.py:40) ----------------- */

/* Output length of analysis: floor((T + 2*((K-1)/2) - K) / N) + 1 with K = taps + 1 (pqmf.py:49-50). */
IAS_API int ias_pqmf_out_len(int T, int N, int K);

/* PQMF.analysis / forward (pqmf.py:46-50): out[b][k][n] = sum_j H[k][j] * x[b][n*N + j - (K-1)/2].
 * H is the module buffer H[:,0,:] = [N][K]; H_host is the same values in host memory (fast paths pass the taps
 * as kernel arguments); H_dev is used when H_host is NULL or the shape has no specialised kernel.
 * proto_host[K] / mod_host[N][2N] (both host, both optional) are the cosine-modulated factorisation of H,
 * H[k][j] == proto[j] * mod[k][j % 2N]: when the caller knows H is the filter PQMF.__init__ designs (pqmf.py:18-30)
 * it passes them and the kernel runs the polyphase form (63 + 2N^2 instead of 63N multiply-adds per time step);
 * with NULL the direct form is used, valid for any H (e.g. taps loaded from a checkpoint).
 * The fast kernels rebuild part of the modulation from PQMF.__init__'s own formula (N >= 8 never read mod_host), so a
 * non-NULL factorisation is honoured only if it reproduces the H_host that is passed with that design (centre
 * (taps-1)/2, phase (-1)^k pi/4; checked on the host, cached); any other cosine-modulated bank -- e.g. the textbook
 * taps/2 centring of the reference's TODO at pqmf.py:26 -- runs the direct form.
 * row_scale[B] (may be NULL) multiplies row b of x, so normalize_if_clipping can be folded in (scale = 1/peak). */
IAS_API int ias_pqmf_analysis(const float* x, const float* H_dev, const float* H_host, const float* proto_host,
                      const float* mod_host, const float* row_scale, float* out, int B, int T, int N, int K,
                      ias_stream_t stream);

/* PQMF.analysis fused with the image preprocessing AudioEmbedding._preprocess applies to the bands
 * (audioembed.py:38-49): z.reshape(-1,3,240,245) is a view of out[B][N][L]; img_preprocess is
 * torchvision.transforms.Normalize(mean, std) (vicreg_audio_params.py:60-62), i.e. out[b][k][n] =
 * (band[b][k][n] - mean[k]) / std[k] in fp32, applied in the store epilogue so the bands make one trip to HBM.
 * mean_host[N], std_host[N] are host arrays; norm_dev = device [mean[N] | std[N]] is only read by the generic kernel
 * (shapes without a specialised one) and may be NULL otherwise. */
IAS_API int ias_pqmf_analysis_image(const float* x, const float* H_dev, const float* H_host, const float* proto_host,
                            const float* mod_host, const float* row_scale, const float* mean_host,
                            const float* std_host, const float* norm_dev, float* out, int B, int T, int N, int K,
                            ias_stream_t stream);

/* PQMF.analysis plus pooled band magnitudes -- harness bridge of SURVEY.md 8(d), not a reference surface:
 * feat[b][i] = mean |bands_flat[b][floor(i S/P) .. ceil((i+1) S/P))| with S = N*L, i.e.
 * torch.nn.functional.adaptive_avg_pool1d(out.abs().reshape(B,1,N*L), P), accumulated by the analysis CTAs while the
 * band values are in registers: every warp adds its |values| in 2^-22 fixed point (one redux.sync) into 64-bit bin
 * accumulators in `workspace` (integer adds: bit-identical from run to run whatever the arrival order), a small
 * finalize kernel converts and divides -- so the bands are written once and not read back.  Resolution 2.4e-7 per
 * thread sum; a thread's sum over its steps saturates at 32 (never reached for audio within [-1, 1]).  Requires a specialised kernel (N in {2,3,4,8,16}, K = 63, H_host given)
 * and bins wider than its CTA tile; returns IAS_ERR_UNSUPPORTED otherwise (callers then pool with ias_abs_avg_pool).
 * workspace >= ias_pqmf_pool_workspace_bytes(B, T, N, K) bytes of device memory. */
IAS_API size_t ias_pqmf_pool_workspace_bytes(int B, int T, int N, int K);
IAS_API int ias_pqmf_analysis_pooled(const float* x, const float* H_dev, const float* H_host, const float* proto_host,
                             const float* mod_host, const float* row_scale, float* out, float* feat, int P,
                             void* workspace, size_t workspace_bytes, int B, int T, int N, int K, ias_stream_t stream);

/* PQMF.synthesis (pqmf.py:52-55): zero-stuff by N with gain N, then the N->1 FIR G = [N][K] (buffer G[0]).
 * y[b][t], t < L*N.  proto_host[K] (host, optional) is the same signed prototype as in ias_pqmf_analysis: the caller
 * passes it when G is the filter PQMF.__init__ designs (pqmf.py:18-30), and N = 2, 3, 4, 8, 16 then run the
 * cosine-modulated form (N -> 2N modulation per time step + 63 multiply-adds, instead of 63 N; packed fp32 for
 * N = 3, 4, 8, 16); NULL = direct form, valid for any G.
 * As for the analysis, proto_host is honoured only when G_host is the PQMF.__init__ design it factorises.
 * Tuning / reference switches read from the environment at call time (results are bit-identical either way):
 * IAS_PQMF_SYNTH_PACKED=0 (N = 3, 4: scalar FIR phase), IAS_PQMF_SYNTH_CM2=0 (N = 8, 16: one time step per thread),
 * IAS_PQMF_SYNTH_Q=<steps per thread>, IAS_PQMF_SYNTH_DIRECT=1 (N <= 4: direct form). */
IAS_API int ias_pqmf_synthesis(const float* z, const float* G_dev, const float* G_host, const float* proto_host, float* y,
                       int B, int L, int N, int K, ias_stream_t stream);

/* ---- VICReg loss: vicreg.VICReg.loss / off_diagonal (vicreg.py:35-58,73-76) ------------------------------ */

IAS_API size_t ias_vicreg_workspace_bytes(int B, int D);

/* x, y: [B][D] (the gathered batch when distributed).  Rows [local_row0, local_row0 + B_local) are this rank's own
 * and are the only ones in the invariance term (vicreg.py:36 precedes the gather at vicreg.py:38-39).
 * cfg_batch_size is cfg.vicreg.batch_size (covariance divisor, vicreg.py:47-48), embeddim is cfg.embeddim
 * (vicreg.py:49).  out4 (device) = {loss, repr_loss, std_loss, cov_loss}.  The workspace keeps what
 * ias_vicreg_loss_backward needs until the next forward call on it. */
IAS_API int ias_vicreg_loss(const float* x, const float* y, int B, int local_row0, int B_local, int cfg_batch_size, int D,
                    int embeddim, float sim_coeff, float std_coeff, float cov_coeff, float* out4, void* workspace,
                    size_t workspace_bytes, ias_stream_t stream);

/* d out4 . gout4 / d x, d y for the same arguments as the preceding ias_vicreg_loss on `workspace`.
 * gout4 (device) holds the upstream gradients of {loss, repr, std, cov}.  gx, gy: [B][D]; rows outside the local
 * range receive only the std/cov contributions (FullGatherLayer.backward then reduce-scatters them). */
IAS_API int ias_vicreg_loss_backward(const float* x, const float* y, int B, int local_row0, int B_local, int cfg_batch_size,
                             int D, int embeddim, float sim_coeff, float std_coeff, float cov_coeff,
                             const float* gout4, float* gx, float* gy, void* workspace, size_t workspace_bytes,
                             ias_stream_t stream);

/* Fused embedding all-gather + loss (replaces FullGatherLayer + loss, vicreg.py:36-58 with 38-39 live).
 * x_peers_host[q] / y_peers_host[q] (host arrays of `world` DEVICE pointers, rank order) address rank q's
 * [B_local][D] embeddings in peer-accessible memory (e.g. torch symmetric memory): the column-statistics kernel reads
 * them over NVLink itself and keeps a local copy of the gathered batch in the workspace -- there is no separate
 * all-gather launch.  The caller must have made every rank's buffer visible (a barrier on the stream) before the call
 * and must not overwrite its own buffer until every rank's call has completed.  Invariance term: rows of `rank`. */
IAS_API size_t ias_vicreg_gather_workspace_bytes(int world, int B_local, int D);
IAS_API int ias_vicreg_loss_gather(const float* const* x_peers_host, const float* const* y_peers_host, int world, int rank,
                           int B_local, int cfg_batch_size, int D, int embeddim, float sim_coeff, float std_coeff,
                           float cov_coeff, float* out4, void* workspace, size_t workspace_bytes, ias_stream_t stream);
/* Gradient w.r.t. this rank's own rows, summed over all ranks' losses (what FullGatherLayer.backward delivers): the
 * std/cov part is identical on every rank, so it is `world` times the own-row slice and needs no communication.
 * gx_local, gy_local: [B_local][D].  Uses the workspace of the preceding ias_vicreg_loss_gather. */
IAS_API int ias_vicreg_loss_gather_backward(int world, int rank, int B_local, int cfg_batch_size, int D, int embeddim,
                                    float sim_coeff, float std_coeff, float cov_coeff, const float* gout4,
                                    float* gx_local, float* gy_local, void* workspace, size_t workspace_bytes,
                                    ias_stream_t stream);

/* Statistics exchange (the default multi-GPU route): the loss over the global batch without gathering the
 * embeddings.  Each rank reduces its own [B_local][D] rows with the single-GPU kernels (local mean, locally-centred
 * second moments and tcgen05 Gram), pushes that 4*Dp + 2*ntiles*128*128-float summary (0.39 MB at D = 256) into every
 * rank's inbox with plain stores over NVLink, raises a flag there, and combines the `world` summaries it received with
 * the exact pooled formulas  mu = sum_q B_q mu_q / B,  G = sum_q [G_q + B_q (mu_q - mu)(mu_q - mu)^T]  -- the same
 * numbers (to fp32 rounding) as vicreg.py:40-51 on the rank-ordered concatenation that FullGatherLayer (vicreg.py:38-39,
 * 79-95) would produce, with 1/W of the arithmetic per rank and no collective or barrier launch.
 * buffers_host[q] (host array of `world` DEVICE pointers, rank order) is rank q's exchange buffer of
 * ias_vicreg_stats_buffer_bytes(world, D) bytes in peer-accessible memory (e.g. torch symmetric memory), mapped into
 * this process; the caller zero-fills its own buffer once, sets the int at byte offset 128 (the step counter) to 1 and
 * synchronises all ranks before the first call.  Every rank must make the same sequence of calls (equal B_local).
 * A rank whose peers never arrive traps after 30 s instead of hanging.  x, y: this rank's rows; invariance term on
 * them (vicreg.py:36).  workspace: ias_vicreg_workspace_bytes(B_local, D), kept for the backward. */
IAS_API size_t ias_vicreg_stats_buffer_bytes(int world, int D);
IAS_API int ias_vicreg_loss_stats(const float* x, const float* y, float* const* buffers_host, int world, int rank,
                          int B_local, int cfg_batch_size, int D, int embeddim, float sim_coeff, float std_coeff,
                          float cov_coeff, float* out4, void* workspace, size_t workspace_bytes, ias_stream_t stream);
/* The same call split at the exchange, so a caller can put other work between sending its summary and needing the
 * peers': IAS_STATS_STAGE_PUBLISH runs the local reduction and pushes the summary (out4 unused);
 * IAS_STATS_STAGE_COMBINE waits for every rank's summary of this step, combines and writes out4.  Both bits =
 * ias_vicreg_loss_stats.  (tools/stats_emulate.py uses the split to run W emulated ranks on one GPU in two sweeps.) */
#define IAS_STATS_STAGE_PUBLISH 1
#define IAS_STATS_STAGE_COMBINE 2
IAS_API int ias_vicreg_loss_stats_stages(const float* x, const float* y, float* const* buffers_host, int world, int rank,
                                 int B_local, int cfg_batch_size, int D, int embeddim, float sim_coeff,
                                 float std_coeff, float cov_coeff, float* out4, void* workspace,
                                 size_t workspace_bytes, int stages, ias_stream_t stream);
/* Gradient w.r.t. this rank's rows summed over all ranks' losses (FullGatherLayer.backward, vicreg.py:92-95): the
 * std/cov terms are identical on every rank, so it is `world` times the own-row slice -- no communication.  Same
 * x, y and workspace as the preceding ias_vicreg_loss_stats. */
IAS_API int ias_vicreg_loss_stats_backward(const float* x, const float* y, int world, int B_local, int cfg_batch_size,
                                   int D, int embeddim, float sim_coeff, float std_coeff, float cov_coeff,
                                   const float* gout4, float* gx_local, float* gy_local, void* workspace,
                                   size_t workspace_bytes, ias_stream_t stream);

/* Test hook: plain CUDA-core Gram of the centred matrix, gram[D][D] = xc^T xc, to cross-check the tcgen05 path. */
IAS_API int ias_vicreg_gram_reference(const float* x, int B, int D, float* gram, void* workspace, size_t workspace_bytes,
                              ias_stream_t stream);
/* Test hook: the tcgen05 Gram alone (same workspace).  Needs D % 128 == 0; gram holds [2][D][D] floats (the hook
 * runs x as both sides; use the first [D][D]). */
IAS_API int ias_vicreg_gram_tc(const float* x, int B, int D, float* gram, void* workspace, size_t workspace_bytes,
                       ias_stream_t stream);

/* ---- Harness utility (not a reference surface) -------------------------------------------------------------
 * out[b][i] = mean |x[b][j]| over torch's adaptive_avg_pool1d bin i of a length-S row, i < P.  Used by the benchmark's
 * stand-in for the out-of-scope backbone between the PQMF bands and the embeddings (SURVEY.md 8d "harness bridge"). */
IAS_API int ias_abs_avg_pool(const float* x, float* out, int B, long long S, int P, ias_stream_t stream);

/* ---- Embedding all-gather: vicreg.FullGatherLayer (vicreg.py:79-95; intended call vicreg.py:38-39) ------- */

/* These five live in libias_comm.so (links NCCL); ias_comm_last_error() is that library's error string. */
IAS_API const char* ias_comm_last_error(void);
IAS_API int ias_comm_unique_id(void* id128_host);
IAS_API int ias_comm_init(const void* id128_host, int rank, int world, void** comm);
IAS_API int ias_comm_destroy(void* comm);
/* forward: all[W][count] <- every rank's local[count] in rank order. */
IAS_API int ias_comm_allgather(void* comm, const float* local, float* all, size_t count, ias_stream_t stream);
/* backward: local[count] <- sum over ranks of their all[W][count], own slice (all-reduce + slice == reduce-scatter). */
IAS_API int ias_comm_reduce_scatter(void* comm, const float* all, float* local, size_t count, ias_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* IAS_B200_H */
