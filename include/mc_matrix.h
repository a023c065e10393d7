/* mc_matrix.h -- the matrix_t data contract of the meagre-crowd plugin boundary.
 *
 * ABI mirror of the reference's src/matrix.h:28-99 (enums :28,44,47,50,60 and
 * struct matrix_t :68-99).  Field order, types and enum values are bit-identical
 * so that a matrix_t built by the reference's own file.c / matrix.c can be
 * handed to solver_*_b200() unchanged (72-byte struct on x86-64).
 * The function set below is the part of the reference's matrix.c API that the
 * dispatcher (src/solvers.c) and the driver (src/meagre-crowd.c) call; it is
 * re-stated in meagre-crowd_b200/csrc/host/mc_matrix.c with O(nz log nz)
 * algorithms because BeBOP (the reference's converter) is not vendored.
 */
#ifndef MC_MATRIX_H
#define MC_MATRIX_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/matrix.h:28 */
enum matrix_base_t { FIRST_INDEX_ZERO = 0, FIRST_INDEX_ONE = 1 };
/* src/matrix.h:44 */
enum matrix_format_t { INVALID = 0, DROW, DCOL, SM_COO, SM_CSC, SM_CSR };
/* src/matrix.h:47 */
enum matrix_symmetry_t { SM_UNSYMMETRIC = 0, SM_SYMMETRIC, SM_SKEW_SYMMETRIC, SM_HERMITIAN };
/* src/matrix.h:50 */
enum matrix_symmetric_storage_t { MC_STORE_BOTH = 0, UPPER_TRIANGULAR, LOWER_TRIANGULAR };
/* src/matrix.h:60 */
enum matrix_data_type_t {
  SM_REAL = 0, REAL_DOUBLE = 0, REAL_SINGLE = 1,
  SM_COMPLEX = 2, COMPLEX_DOUBLE = 2, COMPLEX_SINGLE = 3, SM_PATTERN = 4
};

/* src/matrix.h:68-99.  COO: ii=rows, jj=cols.  CSC: ii=row indices,
 * jj=column pointers (n+1).  CSR: ii=row pointers (m+1), jj=column indices.
 * Dense (DROW/DCOL): ii=jj=NULL and nz == m*n. */
typedef struct matrix_t {
  size_t m;
  size_t n;
  size_t nz;
  enum matrix_base_t base;
  enum matrix_format_t format;
  enum matrix_symmetry_t sym;
  enum matrix_symmetric_storage_t location;
  enum matrix_data_type_t data_type;
  void* dd;
  unsigned int* ii;
  unsigned int* jj;
} matrix_t;

/* src/matrix.h:103-127 */
matrix_t* malloc_matrix( void );
void free_matrix( matrix_t* m );
void clear_matrix( matrix_t* m );
matrix_t* copy_matrix( matrix_t* m );
int cmp_matrix( matrix_t* a, matrix_t* b );
int convert_matrix( matrix_t* m, enum matrix_format_t f, enum matrix_base_t b );
int convert_matrix_symmetry( matrix_t* m, enum matrix_symmetric_storage_t loc );
int detect_matrix_symmetry( matrix_t* m );
int validate_matrix( matrix_t* m );
void printf_matrix( char const* const pre, matrix_t* m );

/* src/file.h: load/save by file extension (.mm only in this build) */
int load_matrix( char* n, matrix_t* A );
int save_matrix( matrix_t* A, char* n );
unsigned int matrix_rows( const matrix_t* const A );

#ifdef __cplusplus
}
#endif
#endif
