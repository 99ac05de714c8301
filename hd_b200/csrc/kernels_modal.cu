// kernels_modal.cu — batched Modal path (model.go:719-851). Placeholder until the kernels land.
#include "modal.hpp"

namespace hdb {

struct ModalState {
    int dummy;
};

int modal_mass_rows(const Symbolic &S, int nMass, const int64_t *massN, std::vector<int32_t> &rows,
                    std::vector<std::vector<int32_t>> *massOf)
{
    std::vector<std::vector<int32_t>> lists(S.dof);
    for (int q = 0; q < nMass; q++) {
        if (massN[q] < 0 || massN[q] >= S.nP) return HDB_ERR_INPUT;
        for (int dir = 0; dir < 2; dir++) lists[massN[q] * 3 + dir].push_back(q);
    }
    rows.clear();
    if (massOf) massOf->clear();
    for (int g = 0; g < S.dof; g++)
        if (!lists[g].empty() && S.xrow[g] >= 0) {
            rows.push_back(S.xrow[g]);
            if (massOf) massOf->push_back(lists[g]);
        }
    return (int)rows.size();
}

int modal_enqueue(const Symbolic &, const DevTopo &, ModalState **, const ModalArgs &, cudaStream_t, std::string &detail,
                  int &launches)
{
    launches = 0;
    detail = "modal kernels not built yet";
    return HDB_ERR_UNSUPPORTED;
}

void modal_state_free(ModalState *s) { delete s; }

} // namespace hdb
