/* mc_solver.h -- plugin table and dispatcher interface of meagre-crowd.
 *
 * ABI mirror of the reference's src/solvers.h:45-101 (solver_state_t and the
 * solver_* dispatch API) and src/solver_lookup.h:46-117 (struct
 * solver_properties_t, capability and multicore bit flags).  Values are
 * bit-identical; the dispatcher itself is re-stated in
 * meagre-crowd_b200/csrc/host/mc_solvers.c following src/solvers.c:35-611.
 */
#ifndef MC_SOLVER_H
#define MC_SOLVER_H

#include "mc_matrix.h"
#include <sys/time.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- perftimer (src/perftimer.h:27-40): the dispatcher ticks
 * "analyze"/"factorize"/"evaluate"/"done" define factor and solve seconds
 * (src/solvers.c:478,492,546,610). */
typedef struct perftimer_tic_t {
  struct timeval now;
  unsigned int depth;
  struct perftimer_tic_t* next;
  char* desc;
} perftimer_tic_t;

typedef struct perftimer_t {
  struct perftimer_tic_t* head;
  struct perftimer_tic_t* tail;
  struct perftimer_t* old;
  unsigned int current_depth;
} perftimer_t;

perftimer_t* perftimer_malloc( void );
void perftimer_free( perftimer_t* pT );
int perftimer_inc( perftimer_t* pT, char const* const s, const size_t n );
void perftimer_restart( perftimer_t** ppT );
void perftimer_adjust_depth( perftimer_t* const h, int d );
double perftimer_wall( perftimer_t const* const pT );
double perftimer_wall_av( perftimer_t const* pT );
double perftimer_delta( perftimer_t const* const pT );
unsigned int perftimer_rounds( perftimer_t const* pT );
void perftimer_printf( perftimer_t const* const pT, const unsigned int d );
void perftimer_printf_csv_header( perftimer_t const* const pT, const unsigned int d );
void perftimer_printf_csv_body( perftimer_t const* const pT, const unsigned int d );
/* extension (not in the reference): milliseconds of the named tic, averaged over rounds; <0 if absent */
double perftimer_tic_ms( perftimer_t const* pT, char const* name );

/* ---- src/solvers.h:45-51 */
typedef struct {
  int solver;
  int mpi_rank;
  int verbosity;
  perftimer_t* timer;
  void* specific;
} solver_state_t;

/* ---- src/solver_lookup.h:46-63 */
struct solver_properties_t {
  char* shortname;
  char* name;
  char* author;
  char* organization;
  char* version;
  char* license;
  char* url;
  void ( *init )( solver_state_t* s );
  void ( *analyze )( solver_state_t*, matrix_t* );
  void ( *factorize )( solver_state_t*, matrix_t* );
  void ( *evaluate )( solver_state_t*, matrix_t*, matrix_t* );
  void ( *finalize )( solver_state_t* s );
  unsigned int capabilities;
  unsigned int multicore;
  char* references;
};

/* ---- capability bits, src/solver_lookup.h:66-107 */
#define SOLVES_BASE_ZERO (1<<0)
#define SOLVES_BASE_ONE  (1<<1)
#define SOLVES_FORMAT_DROW (1<<2)
#define SOLVES_FORMAT_DCOL (1<<3)
#define SOLVES_FORMAT_COO  (1<<4)
#define SOLVES_FORMAT_CSC  (1<<5)
#define SOLVES_FORMAT_CSR  (1<<6)
#define SOLVES_SQUARE_ONLY                          (1<<7)
#define SOLVER_SYM_REQUIRES_DIAGONAL                (1<<8)
#define SOLVES_UNSYMMETRIC                          (1<<9)
#define SOLVES_SYMMETRIC_POSITIVE_DEFINITE_ONLY     (1<<10)
#define SOLVES_SYMMETRIC_UPPER_TRIANGULAR           (1<<11)
#define SOLVES_SYMMETRIC_LOWER_TRIANGULAR           (1<<12)
#define SOLVES_SYMMETRIC_BOTH                       (1<<13)
#define SOLVES_SYMMETRIC (SOLVES_SYMMETRIC_UPPER_TRIANGULAR | SOLVES_SYMMETRIC_LOWER_TRIANGULAR | SOLVES_SYMMETRIC_BOTH)
#define SOLVES_SKEW_SYMMETRIC_UPPER_TRIANGULAR      (1<<14)
#define SOLVES_SKEW_SYMMETRIC_LOWER_TRIANGULAR      (1<<15)
#define SOLVES_SKEW_SYMMETRIC (SOLVES_SKEW_SYMMETRIC_UPPER_TRIANGULAR | SOLVES_SKEW_SYMMETRIC_LOWER_TRIANGULAR)
#define SOLVES_HERMITIAN_SYMMETRIC_UPPER_TRIANGULAR (1<<16)
#define SOLVES_HERMITIAN_SYMMETRIC_LOWER_TRIANGULAR (1<<17)
#define SOLVES_HERMITIAN_SYMMETRIC (SOLVES_HERMITIAN_SYMMETRIC_UPPER_TRIANGULAR | SOLVES_HERMITIAN_SYMMETRIC_LOWER_TRIANGULAR)
#define SOLVES_DATA_TYPE_REAL_DOUBLE    (1<<18)
#define SOLVES_DATA_TYPE_REAL_SINGLE    (1<<19)
#define SOLVES_DATA_TYPE_COMPLEX_DOUBLE (1<<20)
#define SOLVES_DATA_TYPE_COMPLEX_SINGLE (1<<21)
#define SOLVES_RHS_BASE_ZERO (1<<22)
#define SOLVES_RHS_BASE_ONE  (1<<23)
#define SOLVES_RHS_DROW (1<<24)
#define SOLVES_RHS_DCOL (1<<25)
#define SOLVES_RHS_COO  (1<<26)
#define SOLVES_RHS_CSC  (1<<27)
#define SOLVES_RHS_CSR  (1<<28)
#define SOLVES_RHS_VECTOR_ONLY (1<<29)

/* ---- multicore bits, src/solver_lookup.h:111-117 */
#define SOLVER_SINGLE_THREADED_ONLY 0
#define SOLVER_REQUIRES_MPI (1<<0)
#define SOLVER_CAN_USE_MPI  (1<<1)
#define SOLVER_REQUIRES_OMP (1<<2)
#define SOLVER_CAN_USE_OMP  (1<<3)

#define SOLVER_SHORTNAME_MAX_LEN 15

/* the table itself lives in mc_solvers.c; terminated by a {0} row like
 * src/solver_lookup.h:338 */
extern const struct solver_properties_t solver_lookup[];

/* ---- dispatcher API, src/solvers.h:56-101 */
int lookup_solver_by_shortname( const char* s );
const char* solver2str( const int solver );
void printf_solvers( const unsigned int verbosity );
int solver_can_do( const int solver, matrix_t* A, matrix_t* b );
int solver_uses_mpi( const int solver );
int solver_requires_mpi( const int solver );
int solver_uses_omp( const int solver );
int solver_requires_omp( const int solver );
int select_solver( matrix_t* A, matrix_t* b );
void solver( const int solver, const int verbosity, const int mpi_rank, matrix_t* A, matrix_t* b, matrix_t* x );
void solver_solve( solver_state_t* state, matrix_t* A, matrix_t* b, matrix_t* x );
solver_state_t* solver_init( const int solver, const int verbosity, const int mpi_rank, perftimer_t* timer );
void solver_finalize( solver_state_t* p );
void solver_analyze( solver_state_t* p, matrix_t* A );
void solver_factorize( solver_state_t* p, matrix_t* A );
void solver_evaluate( solver_state_t* p, matrix_t* b, matrix_t* x );

/* src/meagre-crowd.c:51-83: element-wise |x-e| <= precision after DCOL conversion */
int results_match( matrix_t* expected_matrix, matrix_t* result_matrix, const double precision );

#ifdef __cplusplus
}
#endif
#endif
