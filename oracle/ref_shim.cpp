// ref_shim.cpp -- extern "C" entry points around the reference's OWN Fastor-free native headers, compiled from where they lie
// under /root/reference (never copied into this repo):
//   Florence/FiniteElements/Assembly/_Assembly_/ComputeSparsityPattern.h   (_ComputeSparsityPattern_, _ComputeDataIndices_)
//   Florence/FiniteElements/Assembly/_Assembly_/SparseAssemblyNative.h     (SparseAssemblyNativeCSR_, ..._RecomputeDataIndex_)
// Built by oracle/Makefile into oracle/_ref/libflorence_ref.so (git-ignored, travels to the GPU box).  TEST INFRASTRUCTURE:
// used by tests/ to check the oracle restatement and the device kernels bit for bit against the reference's real code for the
// integer side of the path (sparsity pattern, slot maps, CSR scatter).  The floating-point LL kernels need Fastor + CBLAS and
// cannot be built.
#include <cstdint>
#include <cstdlib>
#include "ComputeSparsityPattern.h"
#include "SparseAssemblyNative.h"

extern "C" {

__attribute__((visibility("default")))
int ref_compute_sparsity_pattern(const int* elements, const int* idx_start, const int* elem_container, int nvar, int nnode, int nelem,
                                 int nodeperelem, int idx_start_size, int* counts, int* indices) {
    int nnz = 0;
    _ComputeSparsityPattern_(elements, idx_start, elem_container, nvar, nnode, nelem, nodeperelem, idx_start_size, counts, indices, nnz);
    return nnz;
}

__attribute__((visibility("default")))
void ref_compute_data_indices(const int* indices, const int* indptr, int nelem, int nvar, int nodeperelem, const int* elements,
                              const long* sorter, int* data_local_indices, int* data_global_indices) {
    _ComputeDataIndices_(indices, indptr, nelem, nvar, nodeperelem, elements, sorter, data_local_indices, data_global_indices);
}

__attribute__((visibility("default")))
void ref_sparse_assembly_csr(const double* coeff, const int* data_local_indices, const int* data_global_indices, int elem,
                             int local_capacity, double* data) {
    SparseAssemblyNativeCSR_(coeff, data_local_indices, data_global_indices, elem, local_capacity, data);
}

__attribute__((visibility("default")))
void ref_sparse_assembly_csr_search(const double* coeff, int* indices, int* indptr, double* data, int elem, int nvar, int nodeperelem,
                                    const std::uint64_t* sorted_elements, const std::int64_t* sorter) {
    SparseAssemblyNativeCSR_RecomputeDataIndex_(coeff, indices, indptr, data, elem, nvar, nodeperelem, sorted_elements, sorter);
}
}
