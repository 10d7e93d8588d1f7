/*
 * lj_client.c -- a plain C client of the EmDee C ABI (include/emdee.h), in the role of the reference's
 * src/testc.c: builds a simple-cubic Lennard-Jones box, runs a short NVE trajectory through
 * EmDee_boost / EmDee_displace and prints energies. It links unchanged against the CUDA product
 * (emdee_b200/lib/libemdee.so) or against any other library exporting the same ABI:
 *
 *   gcc -O2 -Iinclude examples/lj_client.c -o lj_client -Lemdee_b200/lib -lemdee -lm
 *   LD_LIBRARY_PATH=emdee_b200/lib ./lj_client 4000 50
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "emdee.h"

int main(int argc, char** argv) {
  int N = argc > 1 ? atoi(argv[1]) : 4000;
  int nsteps = argc > 2 ? atoi(argv[2]) : 50;
  const double rho = 0.8, Rc = 2.5, skin = 0.4, dt = 0.004, kT = 1.2;
  const double L = cbrt(N / rho);
  const int nd = (int)ceil(cbrt((double)N));
  double* R = (double*)malloc(3 * (size_t)N * sizeof(double));
  for (int a = 0; a < N; ++a) {
    int k = a / (nd * nd), j = (a - k * nd * nd) / nd, i = a - j * nd - k * nd * nd;
    R[3 * a + 0] = (L / nd) * (i + 0.5);
    R[3 * a + 1] = (L / nd) * (j + 0.5);
    R[3 * a + 2] = (L / nd) * (k + 0.5);
  }
  double box = L;
  tEmDee md = EmDee_system(1, 1, Rc, skin, N, NULL, NULL, NULL);
  EmDee_set_pair_model(md, 1, 1, EmDee_shifted_force(EmDee_pair_lj_cut(1.0, 1.0)), 0.0);
  EmDee_upload(&md, "box", &box);
  EmDee_upload(&md, "coordinates", R);
  EmDee_random_momenta(&md, kT, true, 86245);
  printf("%6d %.12e %.12e %.12e\n", 0, md.Energy.Potential, md.Virial.Total, md.Energy.Potential + md.Kinetic.Total);
  for (int step = 1; step <= nsteps; ++step) {
    md.Options.Compute = (step % 10 == 0);
    EmDee_boost(&md, 1.0, 0.0, 0.5 * dt);
    EmDee_displace(&md, 1.0, 0.0, dt);
    EmDee_boost(&md, 1.0, 0.0, 0.5 * dt);
    if (step % 10 == 0)
      printf("%6d %.12e %.12e %.12e\n", step, md.Energy.Potential, md.Virial.Total,
             md.Energy.Potential + md.Kinetic.Total);
  }
  printf("neighbor list builds = %d\n", md.Builds);
  free(R);
  return 0;
}
