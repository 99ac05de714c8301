// test_hd_api.cpp — the C++ host mirror (hd_api.hpp) on the reference's fixtures, written like the
// reference's own tests (example/example.go:14-49 ConsoleBeam, :138-205 Truss; model_test.go:18-25, 352-405).
// Built and run by tests/test_cpp_host_gpu.py on the GPU box.  Prints "OK" or the first failure.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../hd_b200/csrc/hd_api.hpp"

#define CHECK(c)                                                      \
    do {                                                              \
        if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)

static bool close6(double a, double b) { return std::fabs(a - b) <= 1.01e-5 * std::fabs(b) + 1e-30; }

int main(int argc, char **argv)
{
    { // ConsoleBeam: golden "beam" (example/testdata/beam)
        hd::Model m;
        m.Points = {{0.0, 0.0}, {1.0, 0.0}};
        m.Beams = {{{0, 1}, 12e-4, 120e-6, 2.0e11}};
        m.Supports = {{1, 1, 1}, {0, 0, 0}};
        std::vector<hd::LoadCase> lcs(2);
        lcs[0].LoadNodes = {{1, {0, 2.3, 0}}, {1, {10, 0, 0}}};
        lcs[1].LoadNodes = {{1, {10, 0, 0}}, {1, {0, 2.3, 0}}};
        std::vector<hd::ModalCase> mcs(1);
        mcs[0].ModalMasses = {{1, 10000}};
        std::string report;
        hd::Error err = hd::Run(&report, m, lcs, mcs);
        if (err) std::printf("%s", err.str().c_str());
        CHECK(!err);
        if (argc > 1) { // the caller compares the report with the golden file's digest (example_test.go:12-77)
            FILE *f = std::fopen(argv[1], "wb");
            CHECK(f != nullptr);
            std::fwrite(report.data(), 1, report.size(), f);
            std::fclose(f);
        }
        CHECK(m.Pins.size() == 1); // prepare() fills Pins (model.go:344-346)
        for (auto &lc : lcs) {
            CHECK(close6(lc.PointDisplacementGlobal[1][0], 4.16667e-08));
            CHECK(close6(lc.PointDisplacementGlobal[1][1], 3.19444e-08));
            CHECK(close6(lc.PointDisplacementGlobal[1][2], 4.79167e-08));
            CHECK(close6(lc.BeamForces[0][1], -2.3) && close6(lc.BeamForces[0][2], -2.3) && close6(lc.BeamForces[0][4], 2.3));
            CHECK(close6(lc.Reactions[0][0], -10.0) && close6(lc.Reactions[0][1], -2.3) && close6(lc.Reactions[0][2], -2.3));
        }
        CHECK(mcs[0].Result.size() == 2);
        CHECK(std::fabs(mcs[0].Result[0].Hz - 42.29088) < 1e-5 && std::fabs(mcs[0].Result[1].Hz - 77.21223) < 1e-5);
    }
    { // TestDirectionLoadNode (model_test.go:352-405)
        hd::Model m;
        m.Points = {{0.0, 0.0}, {1.0, 0.0}};
        m.Beams = {{{0, 1}, 12e-4, 120e-6, 2.0e11}};
        m.Supports = {{1, 1, 1}, {0, 0, 0}};
        std::vector<hd::LoadCase> lcs(3);
        lcs[0].LoadNodes = {{1, {100.0, 0, 0}}};
        lcs[1].LoadNodes = {{1, {0, 100.0, 0}}};
        lcs[2].LoadNodes = {{1, {0, 0, 100.0}}};
        std::vector<hd::ModalCase> none;
        CHECK(!hd::Run(m, lcs, none));
        CHECK(lcs[0].PointDisplacementGlobal[1][0] > 0 && lcs[1].PointDisplacementGlobal[1][1] > 0 &&
              lcs[2].PointDisplacementGlobal[1][1] > 0);
    }
    { // linear buckling through the mirror: cantilever column, Euler load pi^2 E J / (4 L^2) (example.go:56-89 family)
        hd::Model m;
        const int n = 8;
        const double L = 3.0, E = 2.0e11, J = 120e-8;
        for (int i = 0; i <= n; i++) { m.Points.push_back({0.0, L * i / n}); m.Supports.push_back({0, 0, 0}); }
        m.Supports[0] = {1, 1, 1};
        for (int i = 0; i < n; i++) m.Beams.push_back({{i, i + 1}, 12e-4, J, E});
        hd::LoadCase lc;
        lc.LoadNodes = {{n, {0.0, -1000.0, 0.0}}};
        lc.AmountLinearBuckling = 2;
        CHECK(!hd::LinearStatic(m, {&lc}));
        CHECK(lc.LinearBucklingResult.size() == 2);
        const double euler = M_PI * M_PI * E * J / (4 * L * L);
        CHECK(std::fabs(lc.LinearBucklingResult[0].Factor * 1000.0 - euler) / euler < 1e-3);
        CHECK(lc.LinearBucklingResult[1].Factor > lc.LinearBucklingResult[0].Factor);
    }
    { // Truss + TestWrongLoad (model_test.go:18-25): a moment on a fully pinned node must be rejected
        hd::Model m;
        m.Points = {{0, 0}, {0, 12}, {4, 0}, {4, 6}, {8, 0}};
        const double E = 2.0e11;
        m.Beams = {{{0, 1}, 40e-4, 1, E}, {{0, 2}, 64e-4, 1, E}, {{0, 3}, 60e-4, 1, E}, {{1, 3}, 60e-4, 1, E},
                   {{2, 3}, 40e-4, 1, E}, {{2, 4}, 64e-4, 1, E}, {{3, 4}, 60e-4, 1, E}};
        m.Supports = {{1, 1, 0}, {0, 0, 0}, {0, 1, 0}, {0, 0, 0}, {0, 1, 0}};
        m.Pins.assign(7, {0, 1, 1, 0, 1, 1});
        hd::LoadCase ok, bad;
        ok.LoadNodes = {{1, {-70000, 0, 0}}, {3, {42000, 0, 0}}};
        bad.LoadNodes = {{1, {-70000, 0, 42.0}}};
        CHECK(!hd::LinearStatic(m, {&ok}));
        CHECK(close6(ok.BeamForces[0][0], 1.05e5) && close6(ok.Reactions[0][1], 9.57501e4));
        if (argc > 2) { // report of a model with pins (Model.String prints Go's true / false), compared by the caller
            std::vector<hd::LoadCase> one(1);
            one[0].LoadNodes = ok.LoadNodes;
            std::vector<hd::ModalCase> none;
            std::string report;
            CHECK(!hd::Run(&report, m, one, none));
            FILE *f = std::fopen(argv[2], "wb");
            CHECK(f != nullptr);
            std::fwrite(report.data(), 1, report.size(), f);
            std::fclose(f);
        }
        hd::Error e = hd::LinearStatic(m, {&ok, &bad});
        CHECK(bool(e));
        CHECK(!ok.PointDisplacementGlobal.empty() && bad.PointDisplacementGlobal.empty()); // earlier case survives
        // no supports at all: singular (model_test.go:101-129)
        hd::Model f;
        f.Points = {{0, 0}, {0, 1}, {1, 1}};
        f.Beams = {{{0, 1}, 12e-4, 120e-6, 2.0e11}, {{1, 2}, 12e-4, 120e-6, 2.0e11}};
        f.Supports = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        hd::LoadCase lc;
        lc.LoadNodes = {{1, {0, 2.3, 0}}};
        CHECK(bool(hd::LinearStatic(f, {&lc})));
        // invalid input: beam index out of range -> check.go rejects, nothing crashes
        f.Beams[1].N = {1, 20};
        CHECK(bool(hd::LinearStatic(f, {&lc})));
    }
    std::printf("OK\n");
    return 0;
}
