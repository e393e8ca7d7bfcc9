/*
 * absinthe_b200 -- C ABI of the B200 launcher for Absinthe's fused/tiled stencil sequences.
 *
 * The reference has no in-process FFI: its hot path is a generated stand-alone C++/OpenMP
 * program per implementation variant (ref template_tiling.cpp:179-397), reached through
 * source generation + process execution + stdout parsing.  This header
 * is the boundary a maintainer binds instead (ctypes stub in INTEGRATION.md): the generated
 * variant is no longer a program but a (descriptor, cubin) pair produced by
 * generate_code("template_tiling.cu", ...), and the functions below replace the pieces of the
 * reference's main():
 *
 *   absinthe_variant_load    <- the per-variant build (ref stencil_generator.py:62-94) + the tile /
 *                               loop tables of main() (ref template_tiling.cpp:223-275)
 *   absinthe_variant_bind    <- the array / view set-up (ref template_tiling.cpp:205-219)
 *   absinthe_variant_apply   <- one pass of HOT LOOP 1 + HOT LOOP 2: every fused group's tile loop
 *                               (ref template_tiling.cpp:301-336) followed by the periodic halo
 *                               update of its outputs (ref :338-368)
 *   absinthe_variant_run     <- the timing loop and its total / halo accumulators
 *                               (ref template_tiling.cpp:285-376; training: template_training.cpp:259-319)
 *   absinthe_make_periodic   <- make_periodic (ref template_tiling.cpp:136-159)
 *   absinthe_fill_rand       <- array() initialisation from std::rand (ref template_tiling.cpp:62-64,209)
 *   absinthe_flush_l2        <- the cache flush between runs (ref template_tiling.cpp:287-297)
 *   absinthe_halo_pack/unpack<- the PUT / WAIT schedule slots the reference never filled in
 *                               (ref stencil_generator.py:23-37), used by the multi-GPU decomposition
 *
 * Conventions: plain pointers and sizes only.  Every function returning int returns 0 on success
 * and a non-zero CUDA / launcher error code otherwise; the message is available from
 * absinthe_last_error() (thread-local).  Fields are fp64, laid out exactly like the reference's
 * array<T,X,Y,Z> (ref template_tiling.cpp:59-79): (X+2HX)(Y+2HY)(Z+2HZ) doubles, i fastest, halos
 * included, no padding.  Device buffers are owned by the caller; the library owns only the
 * inter-group temporaries and scratch it allocates at load time.  One host thread drives one
 * variant at a time; all work is enqueued on the caller's stream.  There is no CPU fallback.
 */
#ifndef ABSINTHE_B200_H
#define ABSINTHE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABSINTHE_ABI_VERSION 1

typedef struct absinthe_variant absinthe_variant;

int absinthe_abi_version(void);
const char* absinthe_last_error(void);

/* number of CUDA devices visible, or a negative error code */
int absinthe_device_count(void);

/* ---- variant life cycle ------------------------------------------------------------------ */

/* descriptor: NUL-terminated text written next to the .cu by generate_code (the .desc file);
 * cubin: image compiled from the generated .cu for sm_100a.  A group line of the descriptor names the
 * kernel and its launch shape (threads, dynamic shared memory, tile count, parameter-block size) and
 * optionally `scratch <bytes per CTA>` and `cluster <CTAs>`: a kernel compiled for thread-block clusters
 * is launched as a whole number of co-resident clusters and is refused by the training loop of
 * absinthe_variant_run. */
int absinthe_variant_load(const char* descriptor, const void* cubin, size_t cubin_bytes,
                          int device, absinthe_variant** out);
void absinthe_variant_free(absinthe_variant* v);

/* dims = {X, Y, Z, HX, HY, HZ} */
int absinthe_variant_info(const absinthe_variant* v, int dims[6], int* n_inputs, int* n_outputs,
                          int* n_groups);
/* name of the idx-th program input (is_output == 0) or output (is_output != 0); order is the
 * sorted order of TILING.INPUTS / TILING.OUTPUTS */
const char* absinthe_variant_field(const absinthe_variant* v, int is_output, int idx);
/* bytes of one field: (X+2HX)(Y+2HY)(Z+2HZ) * 8 */
size_t absinthe_variant_field_bytes(const absinthe_variant* v);
/* kernels launched by one absinthe_variant_apply (group kernels + halo kernels) */
int absinthe_variant_launches(const absinthe_variant* v);
/* 1 if the periodic update is fused into the group kernels (no separate halo kernels) */
int absinthe_variant_fused_halo(const absinthe_variant* v);

/* device pointers, inputs with periodic halos already applied */
int absinthe_variant_bind(absinthe_variant* v, const double* const* d_inputs,
                          double* const* d_outputs);

/* ---- execution --------------------------------------------------------------------------- */

/* one application of the whole sequence, asynchronous on `stream` (a cudaStream_t) */
int absinthe_variant_apply(absinthe_variant* v, void* stream);
/* the same captured once into a CUDA graph and replayed */
int absinthe_variant_apply_graph(absinthe_variant* v, void* stream);
/* only group `group` (1-based, as numbered by the analyzer); with_halo != 0 adds its halo update.
 * Used by the multi-GPU driver to interleave halo exchanges between groups. */
int absinthe_variant_apply_group(absinthe_variant* v, int group, int with_halo, void* stream);
/* only the periodic halo update of the outputs of `group` */
int absinthe_variant_halo_group(absinthe_variant* v, int group, void* stream);
/* names of the outputs of `group` that need a halo update: returns count, names via idx */
int absinthe_variant_group_outputs(const absinthe_variant* v, int group);
const char* absinthe_variant_group_output(const absinthe_variant* v, int group, int idx);
/* device pointer of any field of the variant (inputs, outputs, inter-group temporaries) */
double* absinthe_variant_field_ptr(const absinthe_variant* v, const char* name);

/* The reference's measurement loop: 2*runs applications, every odd one reported.
 * total_ms[r] = group kernels + halo updates, halo_ms[r] = halo updates only (TOTAL includes HALO,
 * ref template_tiling.cpp:368).  flush != 0 evicts L2 before every application (untimed).
 * training != 0 runs the training protocol instead (one tile per SM, no halo update,
 * ref template_training.cpp:276-311) and reports halo_ms = 0. */
int absinthe_variant_run(absinthe_variant* v, void* stream, int runs, int flush, int training,
                         float* total_ms, float* halo_ms);

/* End-to-end with HOST buffers: H2D copy of every input, one application, D2H copy of every
 * output, synchronised on return.  h_inputs must have periodic halos.  A NULL entry of h_inputs
 * skips that copy and uses the field as it is on the device (the program's CONSTANTS -- mask,
 * wgtfac, hhl, ... -- are uploaded once at bind time and not per step). */
int absinthe_variant_apply_host(absinthe_variant* v, const double* const* h_inputs,
                                double* const* h_outputs, void* stream);

/* The same without the final synchronisation: H2D on a private copy stream, the sequence on `stream`,
 * D2H on a second copy stream, so uploads and downloads overlap (full-duplex PCIe).  Up to TWO steps of
 * a variant may be in flight: the upload of step n+1 starts as soon as the kernels of step n have read
 * the device inputs, its kernels start once the download of step n has read the device outputs -- the
 * caller hands step n+1 its own host output buffers.  absinthe_variant_wait_host blocks until the
 * outputs of the OLDEST step not yet waited for are in host memory (no-op when none is pending); a
 * third async call without a wait in between fails with -6. */
int absinthe_variant_apply_host_async(absinthe_variant* v, const double* const* h_inputs,
                                      double* const* h_outputs, void* stream);
int absinthe_variant_wait_host(absinthe_variant* v);

/* Host-upload mode of the two calls above.  enable != 0: only the X*Y*Z interior of every host input
 * crosses the bus (pitched 3-D copy, 8 % fewer bytes at 1024x1024x80) and the periodic halo is rebuilt on
 * the device by the library -- identical to the default as long as the host halos were periodic, which
 * the reference's main() guarantees (template_tiling.cpp:209-221).  Default 0: whole padded fields. */
int absinthe_variant_host_interior(absinthe_variant* v, int enable);

/* ---- field helpers ----------------------------------------------------------------------- */

/* periodic boundary condition on a device field (i, then j, then k faces; corners consistent) */
int absinthe_make_periodic(double* d_field, int X, int Y, int Z, int HX, int HY, int HZ,
                           void* stream);
/* host: glibc srand(seed) when reseed != 0, then dst[n] = (double) rand() for n < count */
void absinthe_fill_rand(double* h_dst, size_t count, unsigned seed, int reseed);
/* write a buffer larger than L2 so the next kernel starts cold */
int absinthe_flush_l2(void* stream);

/* the reference's verifier (ref template_tiling.cpp:378-396) on device fields: over the interior, a point
 * counts as an error when |expected - actual| > max(|expected|, |actual|) * tolerance (the reference
 * uses 1e-7), else as a match.  Synchronises `stream`. */
int absinthe_compare_fields(const double* d_expected, const double* d_actual, int X, int Y, int Z,
                            int HX, int HY, int HZ, double tolerance, unsigned long long* errors,
                            unsigned long long* matches, void* stream);

/* ---- multi-GPU halo faces (domain decomposed along j; rows are contiguous per k-plane) ---- */

/* pack `width` rows starting at padded row `j0` of every k-plane into a dense buffer
 * [planes][width][X+2HX]; unpack is the inverse.  planes = Z+2HZ. */
int absinthe_halo_pack(const double* d_field, double* d_buffer, int X, int Y, int Z, int HX,
                       int HY, int HZ, int j0, int width, void* stream);
int absinthe_halo_unpack(double* d_field, const double* d_buffer, int X, int Y, int Z, int HX,
                         int HY, int HZ, int j0, int width, void* stream);
/* All outputs of `group` at once: pack their first / last HY interior rows into d_low / d_high
 * ([field][plane][HY][X+2HX], one launch; local_faces != 0 first refreshes the i and k periodic faces
 * unless the group kernel already wrote them), and unpack received rows into the low / high j-halo. */
int absinthe_variant_pack_group(absinthe_variant* v, int group, double* d_low, double* d_high,
                                int local_faces, void* stream);
int absinthe_variant_unpack_group(absinthe_variant* v, int group, const double* d_low,
                                  const double* d_high, void* stream);
/* periodic update restricted to the i and k faces (the j faces come from the neighbour ranks) */
int absinthe_make_periodic_ik(double* d_field, int X, int Y, int Z, int HX, int HY, int HZ,
                              void* stream);

/* ---- multi-GPU halo exchange behind the boundary: j-slabs, one rank per GPU, P2P over NVLink -----------
 *
 * The reference reserves a node grid in the variant description -- TILING.NX/NY/NZ with index[] = {0,0,0}
 * (ref template_tiling.cpp:189-203), per-group HALOS, and PUT / WAIT slots in the schedule
 * (ref stencil_generator.py:23-37) -- and never fills them in.  Here: rank r of n owns the j-slab
 * [r*Y, (r+1)*Y) of a periodic X x (n*Y) x Z domain.  Every rank creates an exchange *window* on its
 * GPU (per group: double-buffered receive buffers for the rows of both neighbours, and arrival flags)
 * and exports it as a 64-byte CUDA IPC handle; the host side passes the handles of ranks r-1 and r+1
 * (mod n) to absinthe_comm_connect by whatever channel it has (torch.distributed, MPI, a file).
 *
 *   PUT  (absinthe_variant_put_group):  the first / last HY interior rows of all outputs of the group -- every
 *        plane, full padded width, so i/k halos travel along -- go straight into the neighbours' windows through
 *        peer-mapped pointers, then their arrival flags are raised.  `engine` picks who moves the rows:
 *        ABSINTHE_PUT_KERNEL  one small kernel (lowest latency: for halos the very next kernel reads);
 *        ABSINTHE_PUT_COPY    the copy engines (strided 2-D copies + a stream memory operation for the flag):
 *                             no SM, no register, no shared memory is taken from a compute kernel that runs at
 *                             the same time -- for exchanges that overlap the next group's kernel.
 *   WAIT (absinthe_variant_wait_group): on a stream of the library, wait for both neighbours' flags of the
 *        matching PUT and copy the received rows into the low / high j-halo (same engine as the PUT); `stream`
 *        then waits for that.
 *
 * No pack buffers, no collective library.  Enqueue the PUT after the group's kernel (any stream ordered after
 * it) and the WAIT on the stream of the kernel that reads the halos -- at the latest before the group's kernel
 * runs again: a second PUT of a group without the WAIT of the first fails with -6.  With n = 1 both neighbours
 * are the rank itself and PUT + WAIT equal the periodic update of the j faces. */
#define ABSINTHE_PUT_KERNEL 0
#define ABSINTHE_PUT_COPY 1
#define ABSINTHE_COMM_HANDLE_BYTES 64
int absinthe_comm_create(absinthe_variant* v, int rank, int nranks,
                         unsigned char handle[ABSINTHE_COMM_HANDLE_BYTES]);
int absinthe_comm_connect(absinthe_variant* v, const unsigned char* prev_handle,
                          const unsigned char* next_handle);
int absinthe_variant_put_group(absinthe_variant* v, int group, int engine, void* stream);
int absinthe_variant_wait_group(absinthe_variant* v, int group, void* stream);
/* COMP + PUT in one: the group's kernel itself stores the copies of its edge rows that wrap in j (rows, i/k halos
 * and corners: what its fused periodic update would write into the rank's own j-halo) into the neighbours'
 * windows while it computes; after the kernel only the arrival flags are raised.  The exchange then costs the
 * neighbour's unpack and nothing else on the critical path.  Needs a kernel generated with fused periodic
 * copies (CUDA.FUSE_HALO); any other group falls back to kernel + ABSINTHE_PUT_KERNEL.  WAIT as above. */
int absinthe_variant_apply_group_put(absinthe_variant* v, int group, void* stream);
int absinthe_comm_destroy(absinthe_variant* v);

#ifdef __cplusplus
}
#endif

#endif /* ABSINTHE_B200_H */
