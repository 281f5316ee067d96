/*
 * bdof.h -- C ABI of the B200-native Fresnel multislice engine (libbdof.so).
 *
 * The reference (mdw771/beyond_dof) is pure Python and has no FFI: the hot path is reached
 * through Python functions.  Each entry point below names the reference function whose work
 * it replaces; beyond_dof_b200/ binds them with ctypes and re-exposes the reference's own
 * Python signatures (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 *   bdof_kernel_factors      <- get_kernel                tensorflow_recon/util.py:165-185
 *   bdof_forward             <- multislice_propagate_batch_numpy  tensorflow_recon/npfuncs.py:16-63
 *                               multislice_propagate_batch        tensorflow_recon/util.py:432-508
 *                               multislice_propagate              tensorflow_recon/util.py:360-429
 *   bdof_loss_mag            <- mean((|psi|-|y|)^2)       tensorflow_recon/fullfield.py:115,
 *                                                          cnn_propagator/fullfield.py:106
 *   bdof_adjoint             <- autodiff of the above     tensorflow_recon/fullfield.py:428-435,
 *                                                          cnn_propagator/fullfield.py:329
 *   bdof_pack_db / unpack    <- grid[:, :, :, i] slicing  tensorflow_recon/npfuncs.py:36-37
 *   bdof_patch_gather/scatter<- probe-window cut          tensorflow_recon/ptychography.py:62-76
 *   bdof_cnn_forward         <- multislice_propagate_cnn  cnn_propagator/propagation.py:18-133
 *   bdof_rotate_gather/scatter<- apply_rotation          cnn_propagator/util.py:374-402
 *   bdof_rotate_bilinear(+adj)<- tf.contrib.image.rotate  tensorflow_recon/fullfield.py:96, ptychography.py:39
 *   bdof_adam_step           <- apply_gradient_adam       cnn_propagator/util.py:280-291
 *   bdof_finite_support      <- mask, clip, shrink-wrap   cnn_propagator/fullfield.py:359-368
 *   bdof_forward_host        <- the whole call with HOST buffers in the reference layout
 *
 * Conventions
 *   - Every pointer named d_* is DEVICE memory owned by the caller; h_* is HOST memory.
 *   - Fields are complex64 stored as interleaved (re, im) floats, layout [batch][ny][nx].
 *   - The object is "db": interleaved (delta, beta) float pairs, slice-major
 *     [n_slice][batch][ny][nx][2].  bdof_pack_db converts from the reference's [B,Y,X,Z].
 *   - Return value: 0 = OK, negative = BDOF_E_*, positive = cudaError_t.  A message for the
 *     last failure on the calling thread is available from bdof_last_error().
 *   - A plan is bound to one device and one stream and is not re-entrant; all calls are
 *     stream-asynchronous unless stated.  There is no CPU fallback: without a CUDA device
 *     every compute entry point fails with a cudaError_t.
 */
#ifndef BDOF_H
#define BDOF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bdof_plan bdof_plan;

enum {
    BDOF_OK = 0,
    BDOF_E_BADARG = -1,       /* null pointer / non-positive size */
    BDOF_E_UNSUPPORTED = -2,  /* field size not a supported FFT length (see bdof_size_supported) */
    BDOF_E_STATE = -3,        /* call order (e.g. adjoint before a storing forward) */
    BDOF_E_NOMEM = -4
};

/* plan flags */
enum {
    BDOF_PROPAGATE_LAST = 1u << 0,  /* TF semantics: every slice propagates (util.py:464-483);
                                       default = NumPy semantics, last slice only modulates */
    BDOF_STORE_SLICES   = 1u << 1,  /* keep psi entering every slice for bdof_adjoint */
    BDOF_Z_BROADCAST    = 1u << 2,  /* db holds ONE slice that repeats along z (axially invariant) */
    BDOF_STEPWISE       = 1u << 3   /* the plan is driven slice by slice (bdof_slice_step_seq: tiling, per-slice outputs): one row
                                       pass + one column pass per slice, multiplier tables sequenced for that schedule */
};

/* free-space step after the object (npfuncs.py:43-61) */
enum { BDOF_FREE_NONE = 0, BDOF_FREE_INF = 1, BDOF_FREE_TF = 2 };

int  bdof_version(void);
const char* bdof_last_error(void);
/* number of kernels launched by this library since load (all plans, this process) */
unsigned long long bdof_launch_count(void);

/* FFT lengths: 1 = power of two in [64, 8192] (register-resident line kernels; sweep kernels up to 4096),
 * 2 = any other 2^a 3^b 5^c 7^d <= 2048 (mixed-radix shared-memory passes: the reference's 72 x 72 and 18 x 18
 * ptychography probes, tensorflow_recon/reconstruct_ptycho.py), 0 = unsupported */
int  bdof_size_supported(int n);

/* Separable factors of the reference transfer function, float64 on the host:
 * H[v,u] = phase0 * hy[v] * hx[u], centred, voxel_nm has 3 entries, h*_out are complex128
 * (interleaved re,im doubles) of length ny / nx. */
int  bdof_kernel_factors(double dist_nm, double lmbda_nm, const double* voxel_nm, int ny, int nx,
                         double pi_const, double* hy_out, double* hx_out, double* phase0_out);

int  bdof_plan_create(bdof_plan** out, int ny, int nx, int batch, int n_slice, uint32_t flags,
                      void* cuda_stream);
void bdof_plan_destroy(bdof_plan* p);
/* Re-bind the plan to another stream of its device (every later call is ordered on it).  The caller is responsible for
 * ordering the new stream after work already queued on the old one (the Python host does so with an event). */
int  bdof_plan_set_stream(bdof_plan* p, void* cuda_stream);

/* Per-slice propagator.  h_hy/h_hx: centred complex128 factors (host). phase0 (re,im) is the
 * global phase exp(i k dz) kept out of the fp32 tables and restored analytically.
 * k_dz = 2*pi*dz/lambda is the transmission constant (npfuncs.py:33). */
int  bdof_set_kernel(bdof_plan* p, const double* h_hy, const double* h_hx, double phase0_re,
                     double phase0_im, double k_dz);
/* General (non-separable) centred H[ny][nx], complex128 on the host (util.py:459-461, h=...). */
int  bdof_set_kernel_full(bdof_plan* p, const double* h_H, double k_dz);
/* Free-space step: mode BDOF_FREE_*; factors only for BDOF_FREE_TF. */
int  bdof_set_free_prop(bdof_plan* p, int mode, const double* h_hy, const double* h_hx,
                        double phase0_re, double phase0_im);

/* psi_exit[b] = multislice(db[:, b], probe).  d_probe [ny][nx] complex64 shared by the batch. */
int  bdof_forward(bdof_plan* p, const float* d_db, const float* d_probe, float* d_exit);

/* loss = mean((|exit| - target_mag)^2) over [batch][ny][nx], written to d_loss (one double,
 * device); d_grad_exit (nullable) receives G = dL/dRe + i dL/dIm. loss_scale multiplies both. */
int  bdof_loss_mag(bdof_plan* p, const float* d_exit, const float* d_target_mag, double loss_scale,
                   double* d_loss, float* d_grad_exit);

/* Back-propagate d_grad_exit through the last bdof_forward (plan must have BDOF_STORE_SLICES).
 * d_db_inout holds (delta,beta) on entry and (dL/ddelta, dL/dbeta) on return, in place.
 * With BDOF_Z_BROADCAST the gradient is written to d_grad_out [n_slice][batch][ny][nx][2]
 * instead (required).  d_grad_probe (nullable) [ny][nx] complex64 = sum over batch. */
int  bdof_adjoint(bdof_plan* p, float* d_db_inout, const float* d_grad_exit, float* d_grad_out,
                  float* d_grad_probe);

/* Transmission stash (optional; plans that run the sweep, resident or cluster-resident kernels).  d_stash
 * [n_slice][batch][ny][nx] complex64, caller-owned, or NULL to switch off: bdof_forward (plans with BDOF_STORE_SLICES) then leaves
 * tau_i = exp(i k delta_i - k beta_i) - 1 of every slice there and the next bdof_adjoint lands it instead of (delta, beta) and
 * skips recomputing it (the exponentials are ~12 % of an adjoint kernel).  Intended use: pass the buffer the adjoint writes the gradient to -- d_grad_out, or d_db_inout itself for
 * the in-place adjoint (db then holds t between forward and adjoint; not with BDOF_Z_BROADCAST) -- so that it costs no
 * memory: every tile of t is read just before the gradient of the same tile replaces it. */
int  bdof_plan_set_t_stash(bdof_plan* p, float* d_stash);

/* Window mode (resident small-field kernels only: square 64 x 64 fields, separable kernel): the batch elements are windows of
 * ONE object d_obj_db [n_slice][oy][ox][2] -- the probe windows of the ptychography model (tensorflow_recon/ptychography.py:62-76).
 * d_origin_yx [batch][2] int32 (device) = window origins, pixels outside the object are vacuum.  While set, bdof_forward /
 * bdof_adjoint take the OBJECT as d_db / d_db_inout (read only) and read (delta, beta) straight through the windows (the
 * object slice is L2 resident), so no [n_slice][batch][ny][nx] copy of the object is ever cut; bdof_adjoint writes the
 * per-window gradients to d_grad_out (required; accumulate them with bdof_patch_gather_add), which may also serve as the
 * transmission stash (bdof_plan_set_t_stash).
 * d_origin_yx = NULL switches the mode off.  bdof_plan_is_resident: 0 = per-slice kernels, 1 = the resident kernels (one CTA per
 * 64 x 64 field; window mode available), 2 = the cluster-resident kernels (one cluster of 8 CTAs per 256 x 256 field). */
int  bdof_plan_set_windows(bdof_plan* p, int oy, int ox, const int* d_origin_yx);
int  bdof_plan_is_resident(const bdof_plan* p);

/* Fused gradient accumulation over the fields of a minibatch (the reference sums the gradients of minibatch_size projection
 * angles before one exchange, reconstruct_fullfield.py:30): while on, bdof_adjoint ADDS its gradient to d_grad_out instead of
 * overwriting it -- vector reductions at L2 (red.global.add.v2.f32) from the row kernels, TMA reduce-stores
 * (cp.reduce.async.bulk.tensor) from the column kernels; one contribution per address and call, so the sum over calls is
 * deterministic.  Plans that run the sweep, resident or cluster-resident kernels (not the stepwise / mixed-radix / general-kernel
 * passes, not window mode); the transmission stash must then live in another buffer. */
int  bdof_plan_set_grad_accumulate(bdof_plan* p, int on);

/* layout conversion: reference [B,Y,X,Z] float32 planes <-> slice-major interleaved db */
int  bdof_pack_db(const float* d_delta_byxz, const float* d_beta_byxz, float* d_db, int batch, int ny,
                  int nx, int n_slice, void* cuda_stream);
int  bdof_unpack_db(const float* d_db, float* d_delta_byxz, float* d_beta_byxz, int batch, int ny,
                    int nx, int n_slice, void* cuda_stream);

/* Row-chunked variants for pipelined host transfers: the chunk buffers hold rows [row0, row0 + n_rows) of the
 * total_rows = B*Y rows of the reference-layout arrays ([rows][X][Z], z fastest); d_db is the whole slice-major object. */
int  bdof_pack_db_rows(const float* d_delta_chunk, const float* d_beta_chunk, float* d_db, long long total_rows, long long row0,
                       int n_rows, int nx, int n_slice, void* cuda_stream);
int  bdof_unpack_db_rows(const float* d_db, float* d_delta_chunk, float* d_beta_chunk, long long total_rows, long long row0,
                         int n_rows, int nx, int n_slice, void* cuda_stream);

/* Ptychography windows.  Object db_obj [n_slice][oy][ox][2]; positions (y0,x0) = window origin
 * (may be negative / overhang: zero padding as ptychography.py:45-61); output
 * [n_slice][n_pos][py][px][2].  scatter_add accumulates window gradients back (fp32 atomics). */
int  bdof_patch_gather(const float* d_db_obj, int n_slice, int oy, int ox, const int* d_pos_yx, int n_pos,
                       int py, int px, float* d_db_patches, void* cuda_stream);
int  bdof_patch_scatter_add(const float* d_grad_patches, int n_slice, int oy, int ox, const int* d_pos_yx,
                            int n_pos, int py, int px, float* d_grad_obj, void* cuda_stream);

/* bdof_patch_scatter_add as a deterministic gather: every object pixel sums, in scan-position order, the window pixels that
 * fall on it and adds the sum to d_grad_obj (+=); no atomics, bit-reproducible.  At most 4096
 * positions per call. */
int  bdof_patch_gather_add(const float* d_grad_patches, int n_slice, int oy, int ox, const int* d_pos_yx, int n_pos,
                           int py, int px, float* d_grad_obj, void* cuda_stream);

/* Real-space propagator of cnn_propagator/propagation.py:18-133: h_kernel is the cropped
 * kernel_size^2 complex128 kernel (host).  db as above; exit [batch][ny][nx]. */
int  bdof_cnn_forward(const float* d_db, const float* d_probe, float* d_exit, float* d_work, int batch,
                      int ny, int nx, int n_slice, const double* h_kernel, int kernel_size, double k_dz,
                      void* cuda_stream);

/* Gradient through the real-space propagator (what autograd.grad(calculate_loss) differentiates in cnn_propagator/fullfield.py:
 * 93-121,329 and ptychography.py:30-81,248).  bdof_cnn_forward_store keeps the field entering every slice: d_slices
 * [n_slice + 1][batch][ny][nx] complex64, slices[0] = the probe, slices[n_slice] = the chain output BEFORE the corner-pixel
 * rescaling of propagation.py:109-110 (the host side applies it and its gradient).  bdof_cnn_adjoint back-propagates d_G
 * (gradient w.r.t. slices[n_slice]; [2][batch][ny][nx], second field = work space): d_grad_out [n_slice][batch][ny][nx][2] =
 * (dL/ddelta, dL/dbeta); on return d_G[0] is the gradient w.r.t. the entrance field of every batch element. */
int  bdof_cnn_forward_store(const float* d_db, const float* d_probe, float* d_slices, int batch, int ny, int nx, int n_slice,
                            const double* h_kernel, int kernel_size, double k_dz, void* cuda_stream);
int  bdof_cnn_adjoint(const float* d_db, const float* d_slices, float* d_G, float* d_grad_out, int batch, int ny, int nx,
                      int n_slice, const double* h_kernel, int kernel_size, double k_dz, void* cuda_stream);

/* nearest-neighbour rotation of the object about the y axis, in the (x, z) plane, and its transpose
 * (apply_rotation and the autograd of its fancy index, cnn_propagator/util.py:374-402).  d_obj_db [nz][ny][nx][2];
 * d_lookup_zx [nz][nx][2] int32 = (x_old, z_old) per rotated pixel: the reference's table of one angle
 * (save_rotation_lookup, util.py:295-336; beyond_dof_b200/rotation.py builds it) in slice-major order.  Slice z of the
 * rotated array starts out_slice_stride_px pixels after slice z-1, so it can be a batch element of a plan's db. */
int  bdof_rotate_gather(const float* d_obj_db, const int32_t* d_lookup_zx, float* d_out_db, long long out_slice_stride_px,
                        int ny, int nx, int nz, void* cuda_stream);
int  bdof_rotate_scatter_add(const float* d_grad_rot_db, long long slice_stride_px, const int32_t* d_lookup_zx,
                             float* d_grad_obj_db, int ny, int nx, int nz, void* cuda_stream);

/* The same transpose as a deterministic gather: d_offsets [nz*nx + 1] and d_dest [nz*nx] are the CSR lists of the rotated
 * pixels (z*nx + x) that read from each source pixel (z0*nx + x0); d_grad_obj_db is accumulated into (+=), no atomics. */
int  bdof_rotate_adjoint_csr(const float* d_grad_rot_db, long long slice_stride_px, const int32_t* d_offsets,
                             const int32_t* d_dest, float* d_grad_obj_db, int ny, int nx, int nz, void* cuda_stream);
/* The whole minibatch in one pass: element a's rotated gradient starts a * batch_stride_px pixels after element 0's (the batch
 * axis of a plan's db), its lists are d_offsets[a] / d_dest[a] (HOST arrays of n_angles DEVICE pointers).  The sum over the
 * angles is formed in registers in angle order, so d_grad_obj_db is read and written once per BDOF_ROT_MAX_ANGLES angles;
 * accumulate = 0 overwrites d_grad_obj_db (no zero-fill needed beforehand), 1 adds to it. */
#define BDOF_ROT_MAX_ANGLES 16
int  bdof_rotate_adjoint_csr_batch(const float* d_grad_rot_db, long long slice_stride_px, long long batch_stride_px, int n_angles,
                                   const int32_t* const* d_offsets, const int32_t* const* d_dest, float* d_grad_obj_db,
                                   int accumulate, int ny, int nx, int nz, void* cuda_stream);
/* The same for the slices [z_begin, z_end) of the object gradient only (z_begin a multiple of 32): the back-rotation in z
 * buckets, so that the all-reduce of one bucket can overlap the back-rotation of the next. */
int  bdof_rotate_adjoint_csr_batch_range(const float* d_grad_rot_db, long long slice_stride_px, long long batch_stride_px, int n_angles,
                                         const int32_t* const* d_offsets, const int32_t* const* d_dest, float* d_grad_obj_db,
                                         int accumulate, int ny, int nx, int nz, int z_begin, int z_end, void* cuda_stream);

/* SURVEY 8f-1, TF drivers: tf.contrib.image.rotate(stack([delta, beta], -1), theta, interpolation='BILINEAR')
 * (tensorflow_recon/fullfield.py:96, ptychography.py:39) on the native object d_obj_db [nz][ny][nx][2]: rotation by theta
 * (radians) in the (x, z) plane of every y, about ((nx-1)/2, (nz-1)/2), bilinear taps, zero outside (contrib/image semantics:
 * angles_to_projective_transforms + ProjectiveGenerator).  The adjoint accumulates (+=) the transpose into d_grad_obj_db with
 * fp32 atomics.  TensorFlow 1.x cannot run in the build container, so this pair is checked against a restatement only
 * (parity unpinned by the reference). */
int  bdof_rotate_bilinear(const float* d_obj_db, float* d_out_db, long long out_slice_stride_px, double theta, int ny, int nx, int nz,
                          void* cuda_stream);
int  bdof_rotate_bilinear_adjoint(const float* d_grad_rot_db, long long slice_stride_px, float* d_grad_obj_db, double theta, int ny,
                                  int nx, int nz, void* cuda_stream);

/* Adam update of apply_gradient_adam (cnn_propagator/util.py:280-291), fused over x, g, m, v (fp32, n values):
 * m = (1-b1) g + b1 m; v = (1-b2) g^2 + b2 v; x -= step * (m / (1-b1^(i+1))) / (sqrt(v / (1-b2^(i+1))) + eps). */
int  bdof_adam_step(float* d_x, const float* d_g, float* d_m, float* d_v, long long n, int i_batch, double step_size,
                    double b1, double b2, double eps, void* cuda_stream);

/* the L1 and 3-D total-variation regularisers of the TF driver (tensorflow_recon/fullfield.py:389-396; TV =
 * periodic first differences, L1: util.py:913-923 / cnn_propagator/util.py:61-70) on the native object [nz][ny][nx][2], fused:
 * one pass adds alpha_d sign(delta) + gamma dTV/ddelta and alpha_b sign(beta) to d_grad_db (nullable: value only) and adds
 * alpha_d |delta|_1 + alpha_b |beta|_1 + gamma TV(delta) to *d_loss_inout (double, device).  d_partial_work: 1184 doubles. */
int  bdof_regularizers(const float* d_obj_db, float* d_grad_db, int nz, int ny, int nx, double alpha_d, double alpha_b, double gamma,
                       double* d_loss_inout, double* d_partial_work, void* cuda_stream);

/* finite support, non-negativity and shrink-wrap after every update (cnn_propagator/fullfield.py:359-368):
 * x <- clip(x * mask, 0, inf) on both channels of the interleaved object d_x_db [n_px][2]; d_mask [n_px] fp32 is nullable
 * (clip only); shrink_threshold >= 0 also updates mask <- mask * (delta > shrink_threshold) (the reference uses 1e-15). */
int  bdof_finite_support(float* d_x_db, float* d_mask, long long n_px, double shrink_threshold, void* cuda_stream);

/* End-to-end convenience with HOST buffers in the reference layout: copies delta/beta
 * [B,Y,X,Z] float32 and the probe to the device, packs, runs bdof_forward, copies the exit wave
 * [B,Y,X] complex64 back and synchronises the stream. */
int  bdof_forward_host(bdof_plan* p, const float* h_delta_byxz, const float* h_beta_byxz,
                       const float* h_probe, float* h_exit);

/* One slice of the chain on its own: out = P(in * t(db_slice)) (propagate != 0) or in * t(db_slice).
 * Used by the tiling scheme, which refreshes tile halos between slices.  d_db_slice is
 * [batch][ny][nx][2]; in/out [batch][ny][nx] complex64; out must not alias in.  The global phase
 * exp(i k dz) per propagation is NOT applied (it cancels in every |psi|-based quantity). */
int  bdof_slice_step(bdof_plan* p, const float* d_in, const float* d_db_slice, float* d_out, int propagate);
/* The same for slice `slice_index` of a chain that is stepped in order on a BDOF_STEPWISE plan: uses that slice's entry of the
 * error-feedback multiplier sequence (the product of the fp32 tables applied so far stays within one rounding of the exact
 * power of H), so that 1000 steps do not accumulate the rounding of one fixed table. */
int  bdof_slice_step_seq(bdof_plan* p, const float* d_in, const float* d_db_slice, float* d_out, int propagate, int slice_index);

/* The same (propagating slices only) with the CUT fused into the row pass: the batch elements are windows of a larger pitched
 * complex64 buffer d_buf (a tiling block with its apron); window b's first row starts d_in_offsets[b] elements (device array,
 * int64) into d_buf, consecutive rows are `pitch` elements apart.  d_db_tiles / d_out_tiles keep the window layout
 * [batch][ny][nx].  Window rows of 1024 pixels and more need even offsets and an even pitch. */
int  bdof_slice_step_windows(bdof_plan* p, const float* d_buf, long long pitch, const long long* d_in_offsets, const float* d_db_tiles,
                             float* d_out_tiles, int slice_index);

/* Gradient buckets for the data-parallel all-reduce (Horovod allreduce, tensorflow_recon/fullfield.py:412):
 * the z range is split into n_buckets contiguous buckets counted from the last slice; bdof_adjoint
 * records cuda_events[j] (cudaEvent_t handles owned by the caller) on the plan's stream as soon as
 * bucket j's gradient is final, so a communication stream can reduce it while the sweep continues.
 * n_buckets = 0 disables. */
int  bdof_plan_set_bucket_events(bdof_plan* p, int n_buckets, void** cuda_events);

/* d_out[b] = d_in[b] * d_mult, complex64, d_mult [n_per] shared by the batch (may alias in place): the multiply by
 * get_kernel_ir between the two transforms of the IR free-space step of multislice_propagate (tensorflow_recon/util.py:424-427). */
int  bdof_field_multiply(const float* d_in, const float* d_mult, float* d_out, int batch, long long n_per, void* cuda_stream);

/* The free-space step of the plan on its own (npfuncs.py:43-61): out = free_prop(in), [batch][ny][nx]. */
int  bdof_free_prop(bdof_plan* p, const float* d_in, float* d_out);

/* Adjoint of the free-space step on its own: out = free_prop^H(in) (the first stage of bdof_adjoint). */
int  bdof_free_prop_adjoint(bdof_plan* p, const float* d_in, float* d_out);

/* Bytes of device memory the plan owns (work fields + slice store). */
int  bdof_plan_workspace_bytes(const bdof_plan* p, size_t* bytes_out);

/* Data-parallel gradient exchange over NVLink peer memory with the copy engines (csrc/dpexchange.cu) -- the all-reduce (mean)
 * of the object gradient that Horovod / comm.Allreduce perform in the reference (tensorflow_recon/fullfield.py:412,
 * cnn_propagator/fullfield.py:348-351), without communication kernels on the SMs the sweep kernels own.
 * One context per rank (one process per GPU).  The context OWNS the gradient buffer (it has to be exportable by CUDA IPC):
 * the adjoint must write its gradient at bdof_dp_grad_ptr.  Setup: every rank calls bdof_dp_export, the host side
 * all-gathers the bdof_dp_handle_bytes()-sized handles (any transport) and passes the concatenation [world][handle] to
 * bdof_dp_connect.  Per step: bdof_dp_bucket for every bucket, in the same order on every rank (offset and size in bytes,
 * multiples of 16 * world; ready_event = cudaEvent_t recorded once that part of the gradient is final, e.g. the events of
 * bdof_plan_set_bucket_events), then bdof_dp_finish(stream) which makes `stream` wait for the averaged gradient.
 * grad_bytes must be a multiple of 16 * world. */
typedef struct bdof_dp bdof_dp;
int  bdof_dp_create(bdof_dp** out, int rank, int world, size_t grad_bytes, int n_buckets, int gather_only);
void bdof_dp_destroy(bdof_dp* c);
int  bdof_dp_handle_bytes(void);
int  bdof_dp_export(bdof_dp* c, void* h_handle_out);
int  bdof_dp_connect(bdof_dp* c, const void* h_all_handles);
int  bdof_dp_grad_ptr(bdof_dp* c, void** d_grad_out);
int  bdof_dp_bucket(bdof_dp* c, size_t offset_bytes, size_t n_bytes, void* ready_event);
/* gather half only (contexts created with gather_only = 1, no staging area): shard `rank` of the bucket already holds the
 * reduced values when ready_event fires (in-place NCCL reduce-scatter); it is copied into every peer's gradient, one peer
 * after the other.  Three or more GPUs: concurrent copy-engine transfers to several peers do not add up, and the
 * reduce-scatter needs SM kernels anyway, so the exchange is NCCL reduce-scatter + copy-engine all-gather there. */
int  bdof_dp_gather(bdof_dp* c, size_t offset_bytes, size_t n_bytes, void* ready_event);
int  bdof_dp_finish(bdof_dp* c, void* cuda_stream);

/* Leave n_sms streaming multiprocessors free of the persistent pass kernels (process-wide), so that
 * communication kernels (NCCL all-reduce of the gradient buckets) can run concurrently with the sweep. */
int  bdof_set_sm_reserve(int n_sms);

/* Developer / test hook (host only, no device needed): the per-bin complex gain g_k of one realised fp32 convolution
 * IFFT(h FFT(x)) of length n -- the diagonal part of the deviation of the kernels' transform pair (fp32-rounded twiddle
 * constants) from the exact DFT -- that the multiplier tables are divided by.  gain_out: n complex128 (re, im).  All ones for
 * lengths the model does not cover (mixed radix, 4096, 8192). */
int  bdof_debug_fft_gain(int n, double* gain_out);
/* Host evaluation of the multiply-shift division the back-rotation kernel uses to split a reader-list entry d = z*nx + x
 * (0 <= d < 2^31), for the CPU tests. */
int  bdof_debug_rot_split(int nx, int d, int* z, int* x);

/* Developer hook: device buffer that instrumented builds (-DBDOF_PHASE_TIMING) fill with clock64()
 * phase stamps; ignored by the production build. */
int  bdof_debug_set_buffer(void* d_buf);

/* Tiling-based multislice across GPUs (BASELINE config 5; csrc/tilehalo.cu).  The global field is a gy x gx grid of blocks, one
 * per rank (one process per GPU); rank r owns block (r / gx, r % gx) of by x bx pixels, kept with an apron of `apron` pixels on
 * every side in two exportable buffers [by + 2 apron][bx + 2 apron] complex64 (ping-pong between slices).  Set-up as for
 * bdof_dp_*: every rank exports a handle, the host side all-gathers them and passes the concatenation to bdof_tiles_connect.
 * Per slice: bdof_tiles_cut (local FFT windows [n_tiles][ly][lx] out of buffer `which`; d_origin_yx [n_tiles][2] relative to the
 * block interior, may reach into the apron), bdof_slice_step on the windows, bdof_tiles_paste (the OWNED rectangle of every
 * window, d_own_yxhw [n_tiles][4] = y0, x0, height, width relative to the block interior, into the interior of the other
 * buffer), bdof_tiles_halo_exchange on that buffer: one kernel stores the border strips straight into the aprons of the 8
 * neighbours' buffers over NVLink peer memory (periodic wrap at the global border), flags are raised with stream memory
 * operations and `cuda_stream` waits for the 8 neighbours' flags.  Collective; nothing synchronises with the host.
 * The same cut / paste also serve (delta, beta) blocks (8 bytes per pixel as well). */
typedef struct bdof_tiles bdof_tiles;
int  bdof_tiles_create(bdof_tiles** out, int rank, int gy, int gx, int by, int bx, int apron);
void bdof_tiles_destroy(bdof_tiles* c);
int  bdof_tiles_handle_bytes(void);
int  bdof_tiles_export(bdof_tiles* c, void* h_handle_out);
int  bdof_tiles_connect(bdof_tiles* c, const void* h_all_handles);
int  bdof_tiles_block_ptr(bdof_tiles* c, int which, void** d_out);
int  bdof_tiles_cut(bdof_tiles* c, int which, const int* d_origin_yx, int n_tiles, int ly, int lx, float* d_tiles, void* cuda_stream);
int  bdof_tiles_paste(bdof_tiles* c, int which, const float* d_tiles, const int* d_origin_yx, const int* d_own_yxhw, int n_tiles,
                      int ly, int lx, void* cuda_stream);
int  bdof_tiles_halo_exchange(bdof_tiles* c, int which, void* cuda_stream);

/* In-situ timing: between begin and end every pass-kernel launch of this plan is bracketed by CUDA
 * events on the plan's stream; end() synchronises and returns, per kernel variant, the launch count
 * and the summed device time in ms.  Variants: 0 row conv with transmission, 1 row conv, 2 row conv with
 * the adjoint epilogue, 3 row FFT, 4 row IFFT, 5 col conv (incl. the pipelined column pass), 6 col FFT,
 * 7 col IFFT, 8 col conv with a general 2-D H, 9 (unused: the pipelined column pass reports as 5),
 * 10 sweep kernel forward, 11 sweep kernel adjoint, 12 resident small-field kernel forward,
 * 13 resident small-field kernel adjoint.
 * The event pairs break the programmatic-dependent-launch overlap between consecutive kernels, so the
 * per-launch times are an upper bound (sum > the un-instrumented step): use them for SHARES only. */
int  bdof_profile_begin(bdof_plan* p);
int  bdof_profile_end(bdof_plan* p, int n_variants, int* counts, double* ms_total);

/* Device time of the two halves of the LAST bdof_forward / bdof_adjoint call of this plan, measured by one
 * event pair around each call's whole launch sequence (always on; no per-launch events, so the kernels
 * overlap exactly as in production).  Synchronises the stream.  ms_out[0] = forward, ms_out[1] = adjoint;
 * launches_out[0/1] = kernels launched by each. */
int  bdof_plan_last_times(bdof_plan* p, double* ms_out, int* launches_out);

#ifdef __cplusplus
}
#endif
#endif /* BDOF_H */
