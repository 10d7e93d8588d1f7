// engine.h -- device-resident state and the CUDA hot path behind the EmDee C ABI.
//
// The host shim (abi.cpp) owns model semantics (setup order, mixing, modifiers: reference
// src/EmDeeData.f90:193-264, src/modelClass_*.f90) and hands the engine flat POD tables; the engine
// owns everything that lives in HBM and every kernel (engine.cu). No reference type appears here.
#pragma once

#include <vector>

#include "nb_math.h"

namespace emdee {

// One (itype, jtype) cell of a layer's interaction table (reference pairContainer,
// src/modelClass_pair.f90:66-74, after set_pair_type / mixing / modifier_setup).
struct PairEntry {
  nb::DevModel model;
  double kCoul;
  int coulomb;   // pair%coulomb
  int pad;
};

struct LayerTable {
  std::vector<PairEntry> pair;   // ntypes*ntypes, index = itype*ntypes + jtype (0-based, symmetric)
  nb::DevModel coul;             // me%coul(layer)%model
  bool pairs_exist = false;      // me%pairs_exist(layer)
  bool useInRc = false;          // me%useInRc(layer)
};

struct ForceScalars {
  double Epair = 0, Ecoul = 0, Wpair = 0, Wcoul = 0, Wbody = 0;
};

struct KineticScalars {
  double twoKE[3] = {0, 0, 0};   // sum over free atoms of p^2/m per dimension
};

struct KineticAll {
  double twoKEt[3] = {0, 0, 0};  // free atoms p^2/m + bodies pcm^2/M, per dimension
  double twoKEr[3] = {0, 0, 0};  // bodies I*omega^2, per principal axis
};

// One bonded term (reference tStruct, src/structs.f90:27-30): a bond a0-a1 or an angle a0-a1-a2 (vertex a1), 0-based atoms
struct BondedTerm {
  int a0, a1, a2, kind;   // kind: 0 bond_none, 1 bond_harmonic, 2 angle_none, 3 angle_harmonic
  double p1, p2;          // k, r0 | k, theta0
};

struct BondedScalars {
  double Ebond = 0, Wbond = 0, Eangle = 0, Wangle = 0, Wbody = 0, Ecoul = 0;
};

// Reciprocal-space Ewald solver, as set up on the host (reference cKspaceModel_initialize + kspace_ewald_update)
struct EwaldSetup {
  double alpha = 0, beta = 0;
  int ntk = 0;                    // distinct types of charged atoms
  std::vector<int> atomKType;     // per atom: index among those types, or -1
  std::vector<int> n;             // 3 per wave vector (half space)
  std::vector<double> prefac;     // per wave vector
  std::vector<double> lambda;     // (nlayers, ntk, ntk) Coulomb constants of the type pairs
};

struct EngineStats {
  long long launches = 0, force_launches = 0, build_launches = 0;
  double force_ms = 0, build_ms = 0;
  long long list_entries = 0, interacting = 0;
  int cells_per_dim = 0, device = 0;
};

class Engine {
 public:
  Engine(int natoms, int ntypes, int nlayers, double Rc, double skin, const int* atomType1,
         const double* mass, const double* invMass, const int* atomBody, int nbodies);
  ~Engine();

  // ---- setup-time state --------------------------------------------------------------------
  void set_inner_cutoff(double InRc);                                    // EmDee_layer_based_parameters
  void set_exclusions(const std::vector<int>& first, const std::vector<int>& last,
                      const std::vector<int>& item);                     // 1-based CSR, reference layout
  void set_charges(const double* q);
  void set_interact(const std::vector<char>& interact);                  // ntypes*ntypes
  void set_layer(int layer0, const LayerTable& t);

  // ---- transfers (host pointers; pinned or pageable) ---------------------------------------
  void upload_coordinates(const double* R);
  void upload_momenta(const double* P);
  void upload_forces(int layer0, const double* F);
  void download_coordinates(double* R);
  void download_momenta(double* P);
  void download_forces(int layer0, double* F);

  // ---- hot path ----------------------------------------------------------------------------
  // Neighbor-list maintenance (reference handle_neighbor_lists) + pair loop (compute_pairs).
  // Returns true if the list was rebuilt.
  bool compute_forces(int layer0, bool compute, double Lbox, ForceScalars& out, double& neighbor_seconds);

  // ---- device-resident dynamics for free atoms (reference boost / move / kinetic_energies) --
  void boost(int layer0, double CP, double CF, bool want_kinetic, KineticScalars& ke);
  // announce the kick that follows the next compute_forces of this layer (EmDee_boost with stale forces): it is then
  // launched behind the pair kernel and shares its host wait / reduction; boost() afterwards just returns its sums
  void plan_kick(int layer0, double CP, double CF, bool want_kinetic);
  void displace(double CR, double CP);

  // ---- rigid bodies (reference src/ArBee.f90; device-resident, one thread per body) -------------
  enum BodyItem { BODY_QUATERNION, BODY_QUATMOM, BODY_OMEGA, BODY_RCM, BODY_PCM, BODY_FORCE, BODY_TORQUE, BODY_INERTIA };
  void set_bodies(const std::vector<int>& first, const std::vector<int>& atoms,
                  const std::vector<double>& memberMass);                // CSR of body members (0-based atoms)
  void update_body_frames(double Lbox);                                  // update_rigid_bodies + tBody_update on the uploaded coordinates
  void boost_all(int layer0, double CP, double CF, bool translate, bool rotate, bool want_kinetic, KineticAll& ke);
  void move_all(double CR, double CP, double dt, bool translate, bool rotate, int mode);
  void refresh_member_momenta();                                         // particle_momenta into P (before download_momenta)
  void take_member_momenta(KineticAll& ke);                              // assign_momenta from P (after upload_momenta)
  void download_body(int what, double* out);                             // (width, nbodies), out[b*width + c]
  void upload_body(int what, const double* in);
  void derive_quaternion_momenta();                                      // pi = B(q) 2 I omega
  void shadow_pre(int layer0, double dt, int mode);                      // EmDee_verlet_step pre_force bookkeeping
  void shadow_post(int layer0, double dt, int mode, double& Us, double& Ks_t, double& Ks_r);

  // ---- bonded terms (reference compute_bonds / compute_angles, src/EmDeeData.f90:443-550) --------
  void set_bonded(const std::vector<BondedTerm>& terms);
  void add_bonded(int layer0, double Lbox, bool bonded, bool kspace, BondedScalars& out);   // adds to the layer's forces (after compute_forces)

  // ---- reciprocal space (reference compute_kspace, src/EmDeeData.f90:689-700; src/kspace_ewald.f90:188-314) ----
  void set_ewald(const EwaldSetup& e);
  void add_ewald(int layer0, double Lbox, double& Elong, double& Wbody);   // adds to the layer's forces

  // ---- raw pointers and aliasing (reference EmDee_memory_address / EmDee_share_phase_space) ------
  enum Exposed { EXPOSE_R, EXPOSE_P, EXPOSE_F, EXPOSE_LAYER_F };
  void* expose(int what, int layer0);            // host-dereferenceable pointer to the library's own array
  void share_phase_space(Engine& keep);          // this system gives up R, P, body state for keep's

  // ---- extensions --------------------------------------------------------------------------
  long long pair_count();
  void update_list_stats(int layer0, double Lbox);   // fills list_entries / interacting
  long long download_pairs(int* pairs, long long capacity);
  // Pair-distance histogram over the resident neighbor list (reference EmDee_rdf, src/EmDeeCode.f90:1346-1388).
  // pairSym[it*ntypes + jt] = 0 (pair not wanted) or 1-based packed symmetric index; counts[(sym-1)*bins + bin]
  // comes back with every pair counted ONCE.
  void rdf(double Lbox, int bins, double Rc2_scaled, double bins_by_Rc_scaled, const std::vector<unsigned short>& pairSym,
           int nsym, std::vector<long long>& counts);
  void set_kernel_timing(bool on) { timing_ = on; }
  void tune(const char* knob, int value);   // "local_io" (several GPUs, see upload_coordinates); developer knobs of tools/force_lab.py: "force_variant", "carveout"
  void synchronize();
  int comm_mode();   // 0 single GPU, 1 NCCL per step, 2 NVLink peer kernels per step (NCCL at rebuilds only)
  void io_bytes(long long& h2d, long long& d2h);   // bytes moved so far by coordinate uploads / force downloads
  void* stream_handle();   // the cudaStream_t every kernel of this system is launched on
  EngineStats stats();
  // accumulated CUDA-event time (ms) and launch count per kernel kind (TIMER_*), when kernel timing is on
  enum { TIMER_FORCE = 0, TIMER_BUILD = 1, TIMER_BOOST = 2, TIMER_DISPLACE = 3, TIMER_REFRESH = 4, TIMER_EXCHANGE = 5, TIMER_BINNING = 6, TIMER_OTHER = 7 };
  void kernel_times(double* ms8, long long* n8);

  // ---- multi-GPU: one rank per GPU, z-slab decomposition (NCCL) ---------------------------------
  void comm_init(int rank, int world, const void* nccl_unique_id);

  struct Impl;

 private:
  void rebuild_list(double Lbox);
  void launch_pair_kernel(int layer0, bool compute, double Lbox, bool speculative);
  void launch_planned_kick(bool speculative);
  void flush_kick();   // executes a deferred kick (see boost)
  void collect_planned_kick();
  int timer_begin(int kind);
  void timer_end(int idx);
  void timer_harvest(int idx);
  Impl* d_;
  bool timing_ = false;
  EngineStats stats_;
};

// rank's cell layers [z0, z1) out of M (host helper, no device work)
void slab_range(int M, int rank, int world, int& z0, int& z1);
// 128-byte NCCL unique id (rank 0 creates it, the launcher broadcasts it)
void comm_unique_id(void* out128);

// DFMA microbenchmark on the current device: sustained FP64 FMA throughput in TFLOP/s.
double measure_fp64_fma_tflops();

// Evaluates one of the kernels' arithmetic helpers on n host inputs (what: 0 fast_rcp, 1 nb::rcp, 2 exp_nonpos, 3 uerfc_c,
// 4 nb::uerfc with the library exp); test hook for the device-only code paths.
void math_probe(int what, int n, const double* in, double* out);

}  // namespace emdee
