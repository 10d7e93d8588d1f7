/*
 * emdee.h -- C interface of libemdee.so (B200-native build of EmDee's nonbonded hot path).
 *
 * This header is the drop-in boundary. Every entry point below replaces a `bind(C)` procedure of
 * the reference library (atoms-ufrj/EmDee, version string "15 Oct 2018"); the reference line that
 * defines each one is cited next to its prototype (paths relative to the reference tree).
 *
 * Layout note. The struct below is the layout the reference LIBRARY implements
 * (src/EmDeeCode.f90:36-61, src/EmDeeData.f90:40-64): 240 bytes, `Data` at offset 216, and NO
 * `AutoForceCompute` member in Options. The header shipped with the reference
 * (src/emdee_header.h:1-40) is out of sync with that implementation (it lacks Kinetic.UpToDate
 * and carries an extra Options.AutoForceCompute), so a C client compiled against the shipped
 * header passes a garbage `Data`. This header follows the implementation. See DESIGN.md (Q2).
 *
 * Conventions (unchanged from the reference): atom/type/layer indices are 1-based; arrays are
 * double[N][3] (Fortran (3,N)); option strings are NUL-terminated; errors print
 * "Error in <task>: <msg>." on stderr and exit(1) (src/global.f90:51-56).
 */
#ifndef EMDEE_H
#define EMDEE_H

#ifdef __cplusplus
extern "C" {
#else
#include <stdbool.h>
#endif

typedef struct {              /* src/EmDeeCode.f90:44-49 */
  double Pair;                /* seconds spent in the pair-force path */
  double Motion;              /* seconds spent moving atoms (EmDee_displace) */
  double Neighbor;            /* seconds spent checking / rebuilding the neighbor list */
  double Total;               /* wall clock since the system was created */
} tTime;

typedef struct {              /* src/EmDeeData.f90:40-49 */
  double Potential;           /* sum of all potential-energy terms */
  double Dispersion;          /* pair-model (van der Waals) term */
  double Coulomb;             /* Coulomb-model term */
  double Bond;
  double Angle;
  double Dihedral;
  double ShadowPotential;
  bool   UpToDate;            /* energies correspond to the current configuration */
} tEnergy;

typedef struct {              /* src/EmDeeData.f90:51-59 */
  double Total;               /* translational + rotational kinetic energy */
  double TransPart[3];        /* translational part, per Cartesian direction */
  double Rotational;          /* rotational part (rigid bodies) */
  double RotPart[3];          /* rotational part, per principal axis */
  double ShadowKinetic;
  double ShadowRotational;
  bool   UpToDate;
} tKinetic;

typedef struct {              /* src/EmDeeData.f90:61-64 */
  double Total;               /* internal virial, all terms */
  double Body;                /* rigid-body (constraint) contribution */
} tVirial;

typedef struct {              /* src/EmDeeCode.f90:36-42 */
  bool   Translate;           /* integrate translations in boost/displace */
  bool   Rotate;              /* integrate rigid-body rotations */
  int    RotationMode;        /* free-rotor algorithm selector (0 = exact) */
  bool   AutoBodyUpdate;      /* re-derive body frames when coordinates are uploaded */
  bool   Compute;             /* evaluate energies (false: forces and virial only) */
} tOpts;

typedef struct {              /* src/EmDeeCode.f90:51-61 */
  int      Builds;            /* how many times the neighbor list has been rebuilt */
  tTime    Time;
  tEnergy  Energy;
  tKinetic Kinetic;
  tVirial  Virial;
  int      DoF;               /* degrees of freedom, total */
  int      RotDoF;            /* degrees of freedom, rotational */
  void*    Data;              /* opaque handle of the system state */
  tOpts    Options;
} tEmDee;

/* ---- system construction and state transfer ------------------------------------------------ */

/* src/EmDeeCode.f90:69-208 */
tEmDee EmDee_system( int threads, int layers, double rc, double skin, int N, int* types,
                     double* masses, int* bodies );

/* src/EmDeeCode.f90:212-235 */
void* EmDee_memory_address( tEmDee md, const char* option );

/* src/EmDeeCode.f90:239-269 */
void EmDee_share_phase_space( tEmDee mdkeep, tEmDee* mdlose );

/* src/EmDeeCode.f90:273-305 */
void EmDee_layer_based_parameters( tEmDee md, double InternalRc, int* Apply, int* Bonded );

/* src/EmDeeCode.f90:309-359 */
void EmDee_set_pair_model( tEmDee md, int itype, int jtype, void* model, double kCoul );

/* src/EmDeeCode.f90:363-412 */
void EmDee_set_pair_multimodel( tEmDee md, int itype, int jtype, void* model[], double kCoul[] );

/* src/EmDeeCode.f90:416-444 */
void EmDee_set_kspace_model( tEmDee md, void* model );

/* src/EmDeeCode.f90:448-482 */
void EmDee_set_coul_model( tEmDee md, void* model );

/* src/EmDeeCode.f90:486-520 */
void EmDee_set_coul_multimodel( tEmDee md, void* model[] );

/* src/EmDeeCode.f90:524-570 */
void EmDee_ignore_pair( tEmDee md, int i, int j );

/* src/EmDeeCode.f90:574-655 (harmonic bonds and angles run on the device; dihedrals are stored and excluded only,
   exactly as in the reference, whose EmDee_compute_forces never evaluates them) */
void EmDee_add_bond( tEmDee md, int i, int j, void* model );
void EmDee_add_angle( tEmDee md, int i, int j, int k, void* model );
void EmDee_add_dihedral( tEmDee md, int i, int j, int k, int l, void* model );

/* src/EmDeeCode.f90:659-801 */
void EmDee_download( tEmDee md, const char* option, double* address );

/* src/EmDeeCode.f90:805-925 */
void EmDee_upload( tEmDee* md, const char* option, double* address );

/* src/EmDeeCode.f90:929-946 */
void EmDee_switch_model_layer( tEmDee* md, int layer );

/* src/EmDeeCode.f90:950-1020 */
void EmDee_random_momenta( tEmDee* md, double kT, bool adjust, int seed );

/* src/EmDeeCode.f90:1024-1065 */
void EmDee_boost( tEmDee* md, double lambda, double alpha, double dt );

/* src/EmDeeCode.f90:1069-1103 */
void EmDee_displace( tEmDee* md, double lambda, double alpha, double dt );

/* src/EmDeeCode.f90:1107-1211 */
void EmDee_verlet_step( tEmDee* md, double dt );

/* src/EmDeeCode.f90:1215-1277 -- THE hot entry: neighbor-list maintenance + pair forces */
void EmDee_compute_forces( tEmDee* md );

/* src/EmDeeCode.f90:1281-1395 */
void EmDee_rdf( tEmDee md, int bins, double Rc, int pairs, int itype[], int jtype[], double g[] );

/* ---- model modifiers (src/modelClass_nonbonded.f90:83-241) --------------------------------- */

void* EmDee_shifted( void* model );
void* EmDee_shifted_force( void* model );
void* EmDee_smoothed( void* model, double skin );
void* EmDee_shifted_smoothed( void* model, double skin );
void* EmDee_square_smoothed( void* model, double skin );
void* EmDee_shifted_square_smoothed( void* model, double skin );

/* ---- "none" models -------------------------------------------------------------------------- */

void* EmDee_pair_none( void );      /* src/modelClass_pair.f90:146-150 */
void* EmDee_coul_none( void );      /* src/modelClass_coul.f90:131-136 */
void* EmDee_bond_none( void );      /* src/modelClass_bond.f90 */
void* EmDee_angle_none( void );     /* src/modelClass_angle.f90 */
void* EmDee_dihedral_none( void );  /* src/modelClass_dihedral.f90 */

/* ---- generated model constructors (src/make_models_module.sh:76-137, make_c_header.sh) ----- */

void* EmDee_pair_lj_cut( double epsilon, double sigma );                     /* src/pair_lj_cut.f90:35-69 */
void* EmDee_pair_softcore_cut( double epsilon, double sigma, double lambda );/* src/pair_softcore_cut.f90:39-82 */
void* EmDee_coul_cut( void );                                                /* src/coul_cut.f90:33-53 */
void* EmDee_coul_sf( void );                                                 /* src/coul_sf.f90:32-57 */
void* EmDee_coul_damped( double damp );                                      /* src/coul_damped.f90:33-64 */
void* EmDee_coul_long( void );                                               /* src/coul_long.f90:33-73 */
void* EmDee_coul_damped_smoothed( double damp, double skinWidth );           /* src/coul_damped_smoothed.f90:33-85 */
void* EmDee_coul_damped_square_smoothed( double damp, double skinWidth );    /* src/coul_damped_square_smoothed.f90:33-84 */
void* EmDee_coul_square_smoothed( double skinWidth );                        /* src/coul_square_smoothed.f90:33-76 */
void* EmDee_coul_shifted_square_smoothed( double skinWidth );                /* src/coul_shifted_square_smoothed.f90:33-79 */
void* EmDee_bond_harmonic( double k, double r0 );                            /* src/bond_harmonic.f90 */
void* EmDee_angle_harmonic( double k, double theta0 );                       /* src/angle_harmonic.f90 */
void* EmDee_kspace_ewald( double accuracy );                                 /* src/kspace_ewald.f90:62-78 */

#ifdef __cplusplus
}
#endif

#endif /* EMDEE_H */
