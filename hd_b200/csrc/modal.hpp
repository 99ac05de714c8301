// modal.hpp — host interface of the batched Modal path (kernels in kernels_modal.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "device_topo.hpp"
#include "hdb_internal.hpp"

namespace hdb {

struct ModalState; // device copies of the mass-DOF lists of the last call, scratch matrices

struct ModalArgs {
    int nModels;
    const double *sections; // device [nModels][nB][3]
    int nMass;
    const int64_t *massN;   // host [nMass]
    const double *mass;     // device [nModels][nMass]
    int maxModes;
    int32_t *nModes;        // device [nModels]
    double *hz;             // device [nModels][maxModes]
    double *shapes;         // device [nModels][maxModes][dof] or NULL
    int32_t *status;        // device [nModels]
    double *W, *Ksf, *keScratch;
    int ld, subBatch, smemOptin;
};

// distinct free mass DOFs (X and Y of every mass point that is not eliminated), as x rows,
// ascending DOF; massOf (optional) lists for every row the indices of the masses that add to it.
int modal_mass_rows(const Symbolic &S, int nMass, const int64_t *massN, std::vector<int32_t> &rows,
                    std::vector<std::vector<int32_t>> *massOf);
int modal_enqueue(const Symbolic &S, const DevTopo &D, ModalState **state, const ModalArgs &A, cudaStream_t st,
                  std::string &detail, int &launches);
void modal_state_free(ModalState *s);

} // namespace hdb
