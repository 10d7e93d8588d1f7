/*
 * emdee_ext.h -- extensions that the reference API has no slot for.
 *
 * The reference (atoms-ufrj/EmDee) keeps its neighbor lists private (src/EmDeeData.f90:129) and is a
 * single-process OpenMP library (no device, no ranks). These entry points exist so that
 *   (1) parity tests can compare neighbor-list CONTENTS as pair sets (north_star: "bit-exact as
 *       sorted pair sets") -- they read what src/neighbor_lists.f90:199-300 would have produced;
 *   (2) a benchmark can ask which kernels ran and how long they took on the device;
 *   (3) a multi-process launcher (one rank per GPU) can tell the library its rank layout.
 * None of them is needed by a client that only uses emdee.h.
 */
#ifndef EMDEE_EXT_H
#define EMDEE_EXT_H

#include "emdee.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Number of unordered neighbor pairs currently held (pairs with r^2 < (Rc+skin)^2 at the last
   rebuild, after exclusion / same-body / non-interacting-type filtering). */
long long EmDeeX_pair_count( tEmDee md );

/* Writes the pair list as 0-based atom indices, pairs[2*k] < pairs[2*k+1], in unspecified order.
   `capacity` is the number of pairs the buffer can hold; returns the number written. */
long long EmDeeX_download_pairs( tEmDee md, int* pairs, long long capacity );

/* Release every host and device resource owned by the system (the reference has no destructor). */
void EmDeeX_finalize( tEmDee* md );

/* Implementation tag: "oracle-cpu" for the CPU restatement, "b200-cuda" for the product. */
const char* EmDeeX_backend( void );

#ifndef EMDEE_ORACLE_BUILD
/* ---- product-only: device statistics -------------------------------------------------------- */

typedef struct {
  long long launches;         /* kernels of this library launched since EmDee_system */
  long long force_launches;   /* launches of the pair-force kernel */
  double    force_ms;         /* accumulated device time of the pair-force kernel (CUDA events) */
  long long build_launches;   /* launches of the list-build kernel */
  double    build_ms;         /* accumulated device time of the list-build kernel */
  long long list_entries;     /* neighbor entries held by the device list (full list: 2 per pair) */
  long long interacting;      /* entries with r^2 < Rc^2 at the last force evaluation (if counted) */
  int       cells_per_dim;    /* M of the last rebuild */
  int       device;           /* CUDA device ordinal */
} tEmDeeXStats;

void EmDeeX_stats( tEmDee md, tEmDeeXStats* out );

/* Enable per-kernel CUDA-event timing of the force and build kernels (event pairs in a ring, read back lazily:
   no synchronisation is added to the step). */
void EmDeeX_set_kernel_timing( tEmDee md, int enabled );

/* Accumulated device time (ms) and launch count per kernel kind while kernel timing is enabled:
   [0] pair forces, [1] list build, [2] boost, [3] displace, [4] position refresh, [5] halo/criterion exchange (several GPUs),
   [6] binning + cell sort of a rebuild, [7] unused. Both arrays hold 8 entries. */
void EmDeeX_kernel_times( tEmDee md, double* ms8, long long* n8 );

/* How the ranks of a multi-GPU system talk between two list rebuilds: 0 = single GPU, 1 = NCCL calls every step,
   2 = the library's own kernels over NVLink peer memory (mailboxes + halo stores; NCCL only at rebuilds). */
int EmDeeX_comm_mode( tEmDee md );

/* Bytes moved so far between host and device by EmDee_upload("coordinates") and EmDee_download("forces"). */
void EmDeeX_io_bytes( tEmDee md, long long* h2d, long long* d2h );

/* Block until all queued device work of this system has finished. */
void EmDeeX_synchronize( tEmDee md );

/* Developer knobs of the plain-LJ force kernel ("force_variant": launch shape / cache-policy variant, "carveout":
   preferred shared-memory carveout in percent); used by tools/force_lab.py to time variants on one resident system. */
void EmDeeX_tune( tEmDee md, const char* knob, int value );

/* The CUDA stream (cudaStream_t) all kernels of this system are launched on, so that a client can record its
   own events around a region or order its own work after the library's. */
void* EmDeeX_stream( tEmDee md );

/* DFMA microbenchmark on the current device: measured FP64 FMA throughput in TFLOP/s (the FP64
   roofline denominator; MEASURED_PEAKS.json carries only HBM and bf16 figures). */
double EmDeeX_measure_fp64_tflops( void );

/* Test hook: evaluates one of the kernels' device-only arithmetic helpers on n inputs (what: 0 = reciprocal of the plain-LJ
   pair term, 1 = reciprocal of the model bodies, 2 = exp of a non-positive argument as the typed kernel computes it,
   3 = the reference's erfc(x) (src/math.f90:685-691) as the typed kernel computes it, 4 = the same with the library exp). */
void EmDeeX_math_probe( int what, int n, const double* in, double* out );

/* ---- product-only: multi-GPU (one process per GPU, z-slab decomposition over NCCL) -------------
   Every rank creates the SAME system with the same calls and the same full-size arrays (SPMD). Rank 0
   obtains a 128-byte NCCL id, the launcher broadcasts it (e.g. torch.distributed), and every rank calls
   EmDeeX_comm_init before the first box/coordinates upload. From then on each rank integrates and
   computes forces for the atoms in its cell layers [z0, z1) of the reference's M x M x M grid; ghost
   positions of the two layers above/below come from the neighbor ranks every step (no reverse force
   exchange: the list is full), energies/virial are all-reduced, EmDee_download is a collective that
   reassembles the full array on every rank. */
void EmDeeX_comm_unique_id( char* out128 );
void EmDeeX_comm_init( tEmDee md, int rank, int world, const char* unique_id );
/* cell layers [z0, z1) owned by `rank` out of M (pure host arithmetic) */
void EmDeeX_slab_range( int M, int rank, int world, int* z0, int* z1 );
#endif

#ifdef __cplusplus
}
#endif

#endif /* EMDEE_EXT_H */
