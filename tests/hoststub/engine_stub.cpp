// engine_stub.cpp -- TEST INFRASTRUCTURE: a device-free stand-in for emdee_b200/csrc/engine.cu.
//
// Linking the product's host shim (emdee_b200/csrc/abi.cpp, unchanged) against this stub gives a library
// that runs the reference's host-side semantics -- model setup order, mixing, modifier/cutoff constants,
// interaction masks (reference src/EmDeeData.f90:193-264, src/modelClass_*.f90) -- on a machine without a
// GPU, so the tables the kernels WOULD receive can be compared with the CPU oracle's (tests/test_host_tables.py).
// Every compute entry point is a no-op returning zeros; nothing here is ever shipped.
#include <cstring>
#include <vector>

#include "../../emdee_b200/csrc/engine.h"

namespace emdee {

struct Engine::Impl {
  int N = 0, nt = 1, nlayers = 1;
  std::vector<LayerTable> layers;
  std::vector<char> interact;
  std::vector<double> R, P, F;
};

static Engine::Impl* g_last = nullptr;

Engine::Engine(int natoms, int ntypes, int nlayers, double, double, const int*, const double*, const double*, const int*, int) {
  d_ = new Impl();
  d_->N = natoms;
  d_->nt = ntypes;
  d_->nlayers = nlayers;
  d_->layers.resize(nlayers);
  d_->R.assign(3 * (size_t)natoms, 0.0);
  d_->P.assign(3 * (size_t)natoms, 0.0);
  d_->F.assign(3 * (size_t)natoms * nlayers, 0.0);
  g_last = d_;
}
Engine::~Engine() {
  if (g_last == d_) g_last = nullptr;
  delete d_;
}
void Engine::set_inner_cutoff(double) {}
void Engine::set_exclusions(const std::vector<int>&, const std::vector<int>&, const std::vector<int>&) {}
void Engine::set_charges(const double*) {}
void Engine::set_interact(const std::vector<char>& interact) { d_->interact = interact; }
void Engine::set_layer(int layer0, const LayerTable& t) { d_->layers[layer0] = t; }
void Engine::upload_coordinates(const double* R) { std::memcpy(d_->R.data(), R, d_->R.size() * sizeof(double)); }
void Engine::upload_momenta(const double* P) { std::memcpy(d_->P.data(), P, d_->P.size() * sizeof(double)); }
void Engine::upload_forces(int, const double*) {}
void Engine::download_coordinates(double* R) { std::memcpy(R, d_->R.data(), d_->R.size() * sizeof(double)); }
void Engine::download_momenta(double* P) { std::memcpy(P, d_->P.data(), d_->P.size() * sizeof(double)); }
void Engine::download_forces(int, double* F) { std::memset(F, 0, 3 * (size_t)d_->N * sizeof(double)); }
bool Engine::compute_forces(int, bool, double, ForceScalars& out, double&) {
  out = ForceScalars();
  return false;
}
void Engine::boost(int, double, double, bool, KineticScalars& ke) { ke = KineticScalars(); }
void Engine::displace(double, double) {}
void Engine::set_bodies(const std::vector<int>&, const std::vector<int>&, const std::vector<double>&) {}
void Engine::update_body_frames(double) {}
void Engine::boost_all(int, double, double, bool, bool, bool, KineticAll& ke) { ke = KineticAll(); }
void Engine::move_all(double, double, double, bool, bool, int) {}
void Engine::refresh_member_momenta() {}
void Engine::take_member_momenta(KineticAll& ke) { ke = KineticAll(); }
void Engine::download_body(int, double*) {}
void Engine::upload_body(int, const double*) {}
void Engine::derive_quaternion_momenta() {}
void Engine::set_bonded(const std::vector<BondedTerm>&) {}
void Engine::add_bonded(int, double, bool, bool, BondedScalars& out) { out = BondedScalars(); }
void Engine::set_ewald(const EwaldSetup&) {}
void Engine::add_ewald(int, double, double& e, double& w) { e = w = 0.0; }
void* Engine::expose(int, int) { return nullptr; }
void Engine::share_phase_space(Engine&) {}
void Engine::shadow_pre(int, double, int) {}
void Engine::shadow_post(int, double, int, double& a, double& b, double& c) { a = b = c = 0.0; }
long long Engine::pair_count() { return 0; }
void Engine::update_list_stats(int, double) {}
long long Engine::download_pairs(int*, long long) { return 0; }
void Engine::synchronize() {}
EngineStats Engine::stats() { return stats_; }
void Engine::plan_kick(int, double, double, bool) {}
int Engine::comm_mode() { return 0; }
void Engine::io_bytes(long long& h2d, long long& d2h) { h2d = d2h = 0; }
void Engine::kernel_times(double* ms8, long long* n8) { for (int k = 0; k < 8; ++k) { ms8[k] = 0.0; n8[k] = 0; } }
void Engine::tune(const char*, int) {}
void Engine::rdf(double, int bins, double, double, const std::vector<unsigned short>&, int nsym, std::vector<long long>& counts) {
  counts.assign((size_t)bins * nsym, 0);
}
void* Engine::stream_handle() { return nullptr; }
void Engine::comm_init(int, int, const void*) {}
void slab_range(int M, int rank, int world, int& z0, int& z1) {
  z0 = (int)(((long long)rank * M) / world);
  z1 = (int)(((long long)(rank + 1) * M) / world);
}
void comm_unique_id(void* out128) { std::memset(out128, 0, 128); }
double measure_fp64_fma_tflops() { return 0.0; }
void math_probe(int, int, const double*, double*) {}

}  // namespace emdee

// Same record format as the oracle's EmDeeX_dump_tables (oracle/emdee_oracle.cpp).
extern "C" int EmDeeStub_dump_tables(int layer0, double* out, int cap) {
  using namespace emdee;
  if (g_last == nullptr || layer0 < 0 || layer0 >= g_last->nlayers) return -1;
  const LayerTable& t = g_last->layers[layer0];
  const int nt = g_last->nt;
  const int need = 13 * (nt * nt + 1) + nt * nt + 2;
  if (cap < need) return -need;
  int k = 0;
  auto put = [&](const nb::DevModel& m, double kCoul, double coulomb) {
    const double v[13] = {(double)m.kind, (double)m.modifier, m.eshift, m.fshift, m.Rm, m.factor, m.Rm2fac,
                          m.a, m.b, m.c, m.d, kCoul, coulomb};
    for (double x : v) out[k++] = x;
  };
  for (int i = 0; i < nt; ++i)
    for (int j = 0; j < nt; ++j) {
      const PairEntry& e = t.pair[(size_t)i * nt + j];
      put(e.model, e.coulomb ? e.kCoul : 0.0, (double)e.coulomb);
    }
  put(t.coul, 0.0, 0.0);
  for (int q = 0; q < nt * nt; ++q) out[k++] = g_last->interact.empty() ? 0.0 : (double)(g_last->interact[q] != 0);
  out[k++] = t.pairs_exist ? 1.0 : 0.0;
  out[k++] = t.useInRc ? 1.0 : 0.0;
  return k;
}
