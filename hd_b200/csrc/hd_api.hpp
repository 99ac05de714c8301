// hd_api.hpp — C++ host mirror of hd's exported API over the C ABI (include/hdb200.h). Header-only.
//
// Same names, argument meaning, mutation of arguments and error behaviour as the Go package:
//   hd::Run(m, lcs, mcs)        hd.Run          model.go:926-969
//   hd::LinearStatic(m, lcs)    hd.LinearStatic model.go:358-605   (incl. the linear-buckling block, hdb_buckling_batch)
//   hd::Modal(m, mc)            hd.Modal        model.go:719-851
// Nothing throws for bad input: an Error value is returned (empty == nil) and Run keeps going after a
// failed case (model.go:937-939, 951-954).  The text report (`out io.Writer`) is hd::String(...) and the
// Run overload taking a std::string *out, byte for byte the reference's.  All arithmetic runs in libhdb200.so on
// the GPU.
#pragma once
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
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
    else if (status == HDB_STATUS_NONFINITE && fail >= 0) { upto = (size_t)fail; e.add(name + "Assembly node load: not valid node load"); } // model.go:692-696: earlier cases keep their results
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

// ---- text report: Model.String (model.go:853-924), LoadCase.String (:132-211), ModalCase.String (:235-255) ----
namespace detail {
inline void outf(std::string &o, const char *fmt, ...)
{
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    o += buf;
}
inline void row3(std::string &o, size_t i, const std::array<double, 3> &v) { outf(o, "%5zu %15.5e %15.5e %15.5e\n", i, v[0], v[1], v[2]); }
} // namespace detail

inline std::string String(const Model &m)
{
    using detail::outf;
    std::string o = "\nPoint coordinates:\n";
    outf(o, "%37s %17s\n", "", "   SUPPORT DIRECTION");
    outf(o, "%5s %15s %15s ", "Index", "X, m", "Y, m");
    outf(o, "%5s %5s %5s (0 - free, 1 - fixed)\n", "SX", "SY", "SM");
    for (size_t i = 0; i < m.Points.size(); i++) {
        outf(o, "%5zu %15.5f %15.5f ", i, m.Points[i][0], m.Points[i][1]);
        for (int j = 0; j < 3; j++) outf(o, "%5d ", m.Supports[i][j] ? 1 : 0); // the blank follows all three (model.go:873)
        o += "\n";
    }
    o += "Beam property:\n";
    outf(o, "%5s %15s %15s ", "Index", "Start point", "End point");
    outf(o, "%15s %15s %15s\n", "Area,sq.m", "Moment inertia,m4", "Elasticity,Pa");
    for (size_t i = 0; i < m.Beams.size(); i++) {
        outf(o, "%5zu %15lld %15lld ", i, (long long)m.Beams[i].N[0], (long long)m.Beams[i].N[1]);
        outf(o, "%15.5e %15.5e %15.5e\n", m.Beams[i].A, m.Beams[i].J, m.Beams[i].E);
    }
    bool header = false;
    for (size_t k = 0; k < m.Pins.size(); k++) {
        bool any = false;
        for (int i = 0; i < 6; i++) any = any || m.Pins[k][i];
        if (!any) continue;
        if (!header) {
            o += "Pins of beam in local system coordinate:\n";
            outf(o, "%5s %23s %23s\n", "", "START OF BEAM     ", "END OF BEAM       ");
            outf(o, "%5s %23s %23s\n", "", "------------------", "------------------");
            outf(o, "%5s %7s %7s %7s %7s %7s %7s\n", "Index", "X", "Y", "M", "X", "Y", "M");
            header = true;
        }
        outf(o, "%5zu ", k);
        for (int i = 0; i < 6; i++) outf(o, i < 5 ? "%7s " : "%7s", m.Pins[k][i] ? "true" : "false");
        o += "\n";
    }
    if (!header) o += "All beams haven`t pins\n";
    return o;
}

inline std::string String(const LoadCase &lc)
{
    using detail::outf;
    std::string o = "\nLoad case:\n";
    outf(o, "%5s %15s %15s %15s\n", "Point", "Fx, N", "Fy, N", "M, N*m");
    for (auto &ln : lc.LoadNodes) outf(o, "%5lld %15.5f %15.5f %15.5f\n", (long long)ln.N, ln.Forces[0], ln.Forces[1], ln.Forces[2]);
    if (!lc.PointDisplacementGlobal.empty()) {
        o += "Point displacement in global system coordinate:\n";
        outf(o, "%5s %15s %15s %15s\n", "Point", "DX, m", "DY, m", "Angle, rad.");
        for (size_t i = 0; i < lc.PointDisplacementGlobal.size(); i++) detail::row3(o, i, lc.PointDisplacementGlobal[i]);
    }
    if (!lc.BeamForces.empty()) {
        o += "Local force in beam:\n";
        outf(o, "%5s %45s %45s\n", "", "START OF BEAM     ", "END OF BEAM       ");
        outf(o, "%5s %45s %45s\n", "", "------------------------------------", "------------------------------------");
        outf(o, "%5s %15s %15s %15s %15s %15s %15s\n", "Index", "Fx, N", "Fy, N", "M, N*m", "Fx, N", "Fy, N", "M, N*m");
        for (size_t i = 0; i < lc.BeamForces.size(); i++) {
            outf(o, "%5zu ", i);
            for (int j = 0; j < 6; j++) outf(o, j < 5 ? "%15.5e " : "%15.5e", lc.BeamForces[i][j]);
            o += "\n";
        }
    }
    if (!lc.Reactions.empty()) {
        o += "Reaction on support:\n";
        outf(o, "%5s %15s %15s %15s\n", "Index", "Fx, N", "Fy, N", "M, N*m");
        for (size_t i = 0; i < lc.Reactions.size(); i++) {
            const auto &r = lc.Reactions[i];
            if (r[0] == 0 && r[1] == 0 && r[2] == 0) continue; // points without reaction are not listed (model.go:184-188)
            detail::row3(o, i, r);
        }
    }
    if (!lc.LinearBucklingResult.empty()) {
        o += "\nLinear buckling result:\n";
        for (auto &lbr : lc.LinearBucklingResult) {
            outf(o, "Linear buckling factor: %15.5f\n", lbr.Factor);
            outf(o, "%5s %15s %15s %15s\n", "Point", "X", "Y", "M");
            for (size_t i = 0; i < lbr.PointDisplacementGlobal.size(); i++) detail::row3(o, i, lbr.PointDisplacementGlobal[i]);
        }
    } else if (lc.AmountLinearBuckling == 0) {
        o += "\nLinear buckling result: is not calculated\n";
    } else {
        o += "\nLinear buckling result: haven`t valid data. Probably all beams are tension\n";
    }
    return o;
}

inline std::string String(const ModalCase &mc)
{
    using detail::outf;
    std::string o = "\nModal case:\n";
    outf(o, "%5s %15s\n", "Point", "Mass, N");
    for (auto &mn : mc.ModalMasses) outf(o, "%5lld %15.5f\n", (long long)mn.N, mn.Mass);
    for (auto &mr : mc.Result) {
        outf(o, "Natural frequency : %15.5f Hz\n", mr.Hz);
        outf(o, "%5s %15s %15s %15s\n", "Point", "X", "Y", "M");
        for (size_t i = 0; i < mr.ModalDisplacement.size(); i++) detail::row3(o, i, mr.ModalDisplacement[i]);
    }
    return o;
}

// out == nullptr: no report (hd.Run with a nil writer); otherwise the text hd.Run prints is appended to *out
inline Error Run(std::string *out, Model &m, std::vector<LoadCase> &lcs, std::vector<ModalCase> &mcs)
{
    Error et;
    std::string sink;
    std::string &o = out ? *out : sink;
    o += String(m) + "\n";
    if (!lcs.empty()) {
        std::vector<LoadCase *> arr;
        for (auto &lc : lcs) arr.push_back(&lc);
        o += "Linear Elastic Analysis\n"; // model.go:370-371
        Error e = LinearStatic(m, arr);
        if (e) et.add(e);
        o += "\n\n";
        for (auto &lc : lcs) o += String(lc) + "\n";
    }
    o += "\n\n";
    for (auto &mc : mcs) {
        o += "Modal Analysis\n"; // model.go:729-730
        Error e = Modal(m, mc);
        if (e) {
            et.add(e);
            continue;
        }
        o += "\n\n" + String(mc) + "\n";
    }
    o += "\n\n";
    return et;
}

inline Error Run(Model &m, std::vector<LoadCase> &lcs, std::vector<ModalCase> &mcs) { return Run(nullptr, m, lcs, mcs); }

} // namespace hd
