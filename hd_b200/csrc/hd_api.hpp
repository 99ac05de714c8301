// hd_api.hpp — C++ host mirror of hd's exported API over the C ABI (include/hdb200.h). Header-only.
//
// Same names, argument meaning, mutation of arguments and error behaviour as the Go package:
//   hd::Run(m, lcs, mcs)        hd.Run          model.go:926-969
//   hd::LinearStatic(m, lcs)    hd.LinearStatic model.go:358-605   (buckling block not on the GPU yet)
//   hd::Modal(m, mc)            hd.Modal        model.go:719-851
// Nothing throws for bad input: an Error value is returned (empty == nil) and Run keeps going after a
// failed case (model.go:937-939, 951-954).  The text report (`out io.Writer`) stays host formatting in
// the reference and is not reproduced.  All arithmetic runs in libhdb200.so on the GPU.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/hdb200.h"

namespace hd {

constexpr double Gravity = 9.80665; // model.go:713

struct BeamProp { // model.go:58-77 (same memory layout as hdb_beam)
    std::array<int64_t, 2> N;
    double A, J, E;
};
struct Model { // model.go:18-55
    std::vector<std::array<double, 2>> Points;
    std::vector<BeamProp> Beams;
    std::vector<std::array<uint8_t, 6>> Pins;     // bool per DOF; empty == all rigid
    std::vector<std::array<uint8_t, 3>> Supports;
};
struct LoadNode { // model.go:304-315
    int64_t N;
    std::array<double, 3> Forces;
};
struct LoadCase { // model.go:80-130
    std::vector<LoadNode> LoadNodes;
    std::vector<std::array<double, 3>> PointDisplacementGlobal;
    std::vector<std::array<double, 6>> BeamForces;
    std::vector<std::array<double, 3>> Reactions;
    uint16_t AmountLinearBuckling = 0;
    struct BucklingResult { // model.go:275-291
        double Factor;
        std::vector<std::array<double, 3>> PointDisplacementGlobal;
    };
    std::vector<BucklingResult> LinearBucklingResult;
    void reset() { PointDisplacementGlobal.clear(); BeamForces.clear(); Reactions.clear(); LinearBucklingResult.clear(); } // model.go:214-219
};
struct ModalMass { // model.go:294-301
    int64_t N;
    double Mass;
};
struct ModalResult { // model.go:258-272
    double Hz;
    std::vector<std::array<double, 3>> ModalDisplacement;
};
struct ModalCase { // model.go:222-233
    std::vector<ModalMass> ModalMasses;
    std::vector<ModalResult> Result;
    void reset() { Result.clear(); }
};

struct Error { // stands for Go's error / errors.Tree: empty == nil
    std::vector<std::string> causes;
    explicit operator bool() const { return !causes.empty(); }
    void add(const std::string &s) { causes.push_back(s); }
    void add(const Error &e) { causes.insert(causes.end(), e.causes.begin(), e.causes.end()); }
    std::string str() const {
        std::string s;
        for (auto &c : causes) s += c + "\n";
        return s;
    }
};

namespace detail {
static_assert(sizeof(BeamProp) == sizeof(hdb_beam), "BeamProp must match hdb_beam");
static_assert(sizeof(LoadNode) == sizeof(hdb_load_node), "LoadNode must match hdb_load_node");
static_assert(sizeof(ModalMass) == sizeof(hdb_modal_mass), "ModalMass must match hdb_modal_mass");

inline std::string api_error(int rc, const char *where)
{
    return std::string(where) + ": " + hdb_error_string(rc) + " (" + hdb_last_error_detail() + ")";
}

// prepare(), model.go:336-355: fill Pins when empty (mutates m), run check.go
inline Error prepare(Model &m)
{
    Error e;
    if (m.Pins.empty()) m.Pins.assign(m.Beams.size(), std::array<uint8_t, 6>{});
    char msg[4096] = {0};
    int rc = hdb_check_model((int32_t)m.Points.size(), m.Points.empty() ? nullptr : m.Points[0].data(),
                             (int32_t)m.Beams.size(), reinterpret_cast<const hdb_beam *>(m.Beams.data()),
                             (int32_t)m.Pins.size(), m.Pins.empty() ? nullptr : m.Pins[0].data(),
                             (int32_t)m.Supports.size(), m.Supports.empty() ? nullptr : m.Supports[0].data(), msg, sizeof msg);
    if (rc != HDB_OK) e.add(std::string("check input data: ") + (msg[0] ? msg : hdb_error_string(rc)));
    return e;
}

struct Topo {
    hdb_topology *h = nullptr;
    ~Topo() { hdb_topology_destroy(h); }
};

inline int create(Model &m, Topo &t)
{
    hdb_options opts{};
    opts.device = -1;
    opts.ordering = 2; // automatic: results do not depend on the ordering beyond round-off
    return hdb_topology_create((int32_t)m.Points.size(), m.Points[0].data(), (int32_t)m.Beams.size(),
                               reinterpret_cast<const hdb_beam *>(m.Beams.data()), m.Pins[0].data(), m.Supports[0].data(),
                               &opts, &t.h);
}

inline std::vector<double> sections_of(const Model &m)
{
    std::vector<double> s;
    for (auto &b : m.Beams) { s.push_back(b.A); s.push_back(b.J); s.push_back(b.E); }
    return s;
}
} // namespace detail

inline Error LinearStatic(Model &m, std::vector<LoadCase *> lcs)
{
    const std::string name = "Linear Elastic Analysis: ";
    for (auto *lc : lcs) lc->reset();
    Error et = detail::prepare(m);
    for (auto *lc : lcs) { // check.go:127-163
        char msg[1024] = {0};
        if (hdb_check_loads((int32_t)m.Points.size(), (int32_t)lc->LoadNodes.size(),
                            reinterpret_cast<const hdb_load_node *>(lc->LoadNodes.data()), msg, sizeof msg) != HDB_OK)
            et.add(msg);
    }
    if (et) { Error e; e.add(name + "Prepared input data: " + et.str()); return e; }
    detail::Topo t;
    int rc = detail::create(m, t);
    if (rc != HDB_OK) { Error e; e.add(name + "Assembly node load: " + detail::api_error(rc, "hdb_topology_create")); return e; }
    const size_t nP = m.Points.size(), nB = m.Beams.size(), dof = 3 * nP, nLC = lcs.size();
    std::vector<double> sec = detail::sections_of(m);
    int32_t status = 0, fail = -1;
    if (nLC == 0) { // getK still runs (model.go:410)
        hdb_topology_info info;
        hdb_topology_get_info(t.h, &info);
        std::vector<double> kval((size_t)info.nnzK);
        rc = hdb_assemble_batch(t.h, 1, sec.data(), kval.data(), &status);
        Error e;
        if (rc != HDB_OK) e.add(name + detail::api_error(rc, "hdb_assemble_batch"));
        else if (status != HDB_STATUS_OK) e.add(name + "Assembly node load: not valid stiffness");
        return e;
    }
    std::vector<double> loads(nLC * dof), disp(nLC * dof), forces(nLC * nB * 6), react(nLC * dof);
    for (size_t i = 0; i < nLC; i++) // assemblyNodeLoad, model.go:683-698
        hdb_assemble_node_loads((int32_t)nP, (int32_t)lcs[i]->LoadNodes.size(),
                                reinterpret_cast<const hdb_load_node *>(lcs[i]->LoadNodes.data()), loads.data() + i * dof);
    rc = hdb_static_batch(t.h, 1, sec.data(), (int32_t)nLC, loads.data(), disp.data(), forces.data(), react.data(), &status, &fail);
    Error e;
    if (rc != HDB_OK) { e.add(name + detail::api_error(rc, "hdb_static_batch")); return e; }
    size_t upto = nLC;
    if (status == HDB_STATUS_IGNORE_LOAD) { upto = (size_t)fail; e.add(name + "Warning: List loads on not valid directions"); }
    else if (status == HDB_STATUS_SINGULAR) { upto = 0; e.add(name + "Assembly node load: LU error factorization"); }
    else if (status != HDB_STATUS_OK) { upto = 0; e.add(name + "Assembly node load: not valid value"); }
    for (size_t i = 0; i < upto; i++) {
        auto &lc = *lcs[i];
        lc.PointDisplacementGlobal.resize(nP);
        lc.Reactions.resize(nP);
        lc.BeamForces.resize(nB);
        for (size_t p = 0; p < nP; p++)
            for (int k = 0; k < 3; k++) {
                lc.PointDisplacementGlobal[p][k] = disp[i * dof + 3 * p + k];
                lc.Reactions[p][k] = react[i * dof + 3 * p + k];
            }
        for (size_t b = 0; b < nB; b++)
            for (int k = 0; k < 6; k++) lc.BeamForces[b][k] = forces[(i * nB + b) * 6 + k];
        if (lc.AmountLinearBuckling > 0) { // buckling block, model.go:491-601
            const int amount = lc.AmountLinearBuckling;
            std::vector<double> fac((size_t)amount), shp((size_t)amount * dof);
            int32_t nOut = 0, bst = 0;
            rc = hdb_buckling_batch(t.h, 1, sec.data(), 1, forces.data() + i * nB * 6, amount, amount, &nOut, fac.data(), shp.data(), &bst);
            if (rc != HDB_OK) { e.add(name + detail::api_error(rc, "hdb_buckling_batch")); return e; }
            if (bst != HDB_STATUS_OK) { e.add(name + "Linear buckling: eigen decomposition failed"); return e; }
            for (int r = 0; r < nOut && r < amount; r++) {
                LoadCase::BucklingResult br;
                br.Factor = fac[(size_t)r];
                br.PointDisplacementGlobal.resize(nP);
                for (size_t p = 0; p < nP; p++)
                    for (int k = 0; k < 3; k++) br.PointDisplacementGlobal[p][k] = shp[(size_t)r * dof + 3 * p + k];
                lc.LinearBucklingResult.push_back(br);
            }
        }
    }
    return e;
}

inline Error Modal(Model &m, ModalCase &mc)
{
    const std::string name = "Modal Analysis: ";
    mc.reset();
    Error et = detail::prepare(m);
    char msg[1024] = {0};
    if (hdb_check_modal((int32_t)m.Points.size(), (int32_t)m.Supports.size(), m.Supports.empty() ? nullptr : m.Supports[0].data(),
                        (int32_t)mc.ModalMasses.size(), reinterpret_cast<const hdb_modal_mass *>(mc.ModalMasses.data()), msg,
                        sizeof msg) != HDB_OK)
        et.add(msg);
    if (et) { Error e; e.add(name + et.str()); return e; }
    detail::Topo t;
    int rc = detail::create(m, t);
    if (rc != HDB_OK) { Error e; e.add(name + "Assembly node load: " + detail::api_error(rc, "hdb_topology_create")); return e; }
    std::vector<int64_t> massN;
    std::vector<double> mass;
    for (auto &q : mc.ModalMasses) { massN.push_back(q.N); mass.push_back(q.Mass); }
    const int s = hdb_modal_mode_count(t.h, (int32_t)massN.size(), massN.data());
    Error e;
    if (s <= 0) return e; // empty mass list: no results, no error
    const size_t nP = m.Points.size(), dof = 3 * nP;
    std::vector<double> sec = detail::sections_of(m), hz((size_t)s), shapes((size_t)s * dof);
    int32_t nModes = 0, status = 0;
    rc = hdb_modal_batch(t.h, 1, sec.data(), (int32_t)massN.size(), massN.data(), mass.data(), s, &nModes, hz.data(), shapes.data(), &status);
    if (rc != HDB_OK) { e.add(name + detail::api_error(rc, "hdb_modal_batch")); return e; }
    if (status == HDB_STATUS_EIGEN) { e.add(name + "Eigen factorization is not ok"); return e; }
    if (status != HDB_STATUS_OK) { e.add(name + "Assembly node load: LU error factorization"); return e; }
    for (int k = 0; k < nModes; k++) {
        ModalResult r;
        r.Hz = hz[k];
        r.ModalDisplacement.resize(nP);
        for (size_t p = 0; p < nP; p++)
            for (int c = 0; c < 3; c++) r.ModalDisplacement[p][c] = shapes[(size_t)k * dof + 3 * p + c];
        mc.Result.push_back(r);
    }
    return e;
}

inline Error Run(Model &m, std::vector<LoadCase> &lcs, std::vector<ModalCase> &mcs)
{
    Error et;
    if (!lcs.empty()) {
        std::vector<LoadCase *> arr;
        for (auto &lc : lcs) arr.push_back(&lc);
        Error e = LinearStatic(m, arr);
        if (e) et.add(e);
    }
    for (auto &mc : mcs) {
        Error e = Modal(m, mc);
        if (e) et.add(e);
    }
    return et;
}

} // namespace hd
