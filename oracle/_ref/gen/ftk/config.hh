#ifndef _FTK_CONFIG_HH
#define _FTK_CONFIG_HH

#define FTK_VERSION "0.0.9"

/* #undef FTK_HAVE_ADIOS1 */
/* #undef FTK_HAVE_ADIOS2 */
/* #undef FTK_HAVE_BOOST */
/* #undef FTK_HAVE_CGAL */
/* #undef FTK_HAVE_CUDA */
/* #undef FTK_HAVE_DECAF */
/* #undef FTK_HAVE_GMP */
/* #undef FTK_HAVE_HDF5 */
/* #undef FTK_HAVE_HIPSYCL */
/* #undef FTK_HAVE_SYCL */
/* #undef FTK_HAVE_KOKKOS */
/* #undef FTK_HAVE_LEVELDB */
/* #undef FTK_HAVE_METIS */
/* #undef FTK_HAVE_MPI */
/* #undef FTK_HAVE_MPSOLVE */
/* #undef FTK_HAVE_NETCDF */
/* #undef FTK_HAVE_OPENMP */
/* #undef FTK_HAVE_PARAVIEW */
/* #undef FTK_HAVE_PNETCDF */
/* #undef FTK_HAVE_PNG */
/* #undef FTK_HAVE_PYBIND11 */
/* #undef FTK_HAVE_ROCKSDB */
/* #undef FTK_HAVE_QT5 */
/* #undef FTK_HAVE_QT */
/* #undef FTK_HAVE_TBB */
/* #undef FTK_HAVE_VTK */
/* #undef FTK_HAVE_VTK_JSON */

#define FTK_FP_PRECISION 32768
#define FTK_CP_MAX_NUM_VARS 3

#if FTK_HAVE_MPI
#else
  #define DIY_NO_MPI
#endif

#ifdef __CUDACC__
// #define FTK_NUMERIC_FUNC __device__ __host__
#else
// #define FTK_NUMERIC_FUNC
#define __device__ 
#define __host__ 
#endif

// utilities
#define NC_SAFE_CALL(call) {\
  int retval = call;\
  if (retval != 0) {\
    fprintf(stderr, "[NetCDF Error] %s, in file '%s', line %i.\n", nc_strerror(retval), __FILE__, __LINE__); \
    exit(EXIT_FAILURE); \
  }\
}

#define PNC_SAFE_CALL(call) {\
  int retval = call;\
  if (retval != 0) {\
      fprintf(stderr, "[PNetCDF Error] %s, in file '%s', line %i.\n", ncmpi_strerror(retval), __FILE__, __LINE__); \
      exit(EXIT_FAILURE); \
  }\
}

#endif
