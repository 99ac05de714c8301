// Package hd is the drop-in front of Konstantin8105/hd's linear-analysis path over libhdb200.so: the same exported
// types and the same three entry points,
//
//	func Run(out io.Writer, m *Model, lcs []LoadCase, mcs []ModalCase) error          (model.go:926-969)
//	func LinearStatic(out io.Writer, m *Model, lcs ...*LoadCase) error                (model.go:358-605)
//	func Modal(out io.Writer, m *Model, mc *ModalCase) error                          (model.go:719-851)
//
// plus RunBatch / RunBatchMulti for many models sharing one topology.  Everything the reference delegates to
// Konstantin8105/sparse and gonum/mat on this path runs on the GPU; there is no CPU fallback.
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Go toolchain, DESIGN.md section 1).  The file is the body a maintainer
// drops into the reference tree: there, the original check.go and the String() reports stay as they are and `prepare`
// below calls m.checkInputData() instead of hdb200.CheckModel; tests/cpp/test_hd_api.cpp and hd_b200/api.py replay the
// same call sequence, struct layouts and side effects against the same C ABI.
//
// Side effects kept (SURVEY.md A.9): m.Pins is allocated when empty (model.go:344-346); results are reset at entry
// (model.go:365-367, 726); results of the cases before a failing one survive (model.go:415-437); nothing is printed
// when out == nil except into a discarded buffer (model.go:359-363, 720-724); Run aggregates errors and keeps going
// (model.go:937-939, 951-954); a panic anywhere below becomes an error (model.go:381-385, 740-744).
package hd

import (
	"bytes"
	"errors"
	"fmt"
	"io"
	"runtime/debug"
	"sort"

	"github.com/Konstantin8105/hd/go/hdb200" // in the reference tree: the package that carries the cgo binding
)

// Gravity is the constant of model.go:713.
const Gravity = 9.80665

// Model is hd.Model (model.go:18-55).
type Model struct {
	Points   [][2]float64
	Beams    []BeamProp
	Pins     [][6]bool
	Supports [][3]bool
}

// BeamProp is hd.BeamProp (model.go:58-77): the layout of hdb_beam, passed to C without copying.
type BeamProp struct {
	N       [2]int
	A, J, E float64
}

// LoadNode is hd.LoadNode (model.go:304-315).
type LoadNode struct {
	N      int
	Forces [3]float64
}

// BucklingResult is hd.BucklingResult (model.go:275-291).
type BucklingResult struct {
	Factor                  float64
	PointDisplacementGlobal [][3]float64
}

// LoadCase is hd.LoadCase (model.go:80-130).
type LoadCase struct {
	LoadNodes               []LoadNode
	PointDisplacementGlobal [][3]float64
	BeamForces              [][6]float64
	Reactions               [][3]float64
	AmountLinearBuckling    uint16
	LinearBucklingResult    []BucklingResult
}

func (lc *LoadCase) reset() { // model.go:214-219
	lc.PointDisplacementGlobal = nil
	lc.BeamForces = nil
	lc.Reactions = nil
	lc.LinearBucklingResult = nil
}

// ModalMass is hd.ModalMass (model.go:294-301).
type ModalMass struct {
	N    int
	Mass float64
}

// ModalResult is hd.ModalResult (model.go:258-272).
type ModalResult struct {
	Hz                float64
	ModalDisplacement [][3]float64
}

// ModalCase is hd.ModalCase (model.go:222-228).
type ModalCase struct {
	ModalMasses []ModalMass
	Result      []ModalResult
}

func (mc *ModalCase) reset() { mc.Result = nil } // model.go:231-233

// per-model outcomes of the library (include/hdb200.h HDB_STATUS_*)
const (
	statusOK = iota
	statusNonFinite
	statusSingular
	statusIgnoreLoad
	statusEigen
	statusPattern
)

var errIgnoreLoad = errors.New("Warning: List loads on not valid directions") // model.go:434

func beams(m *Model) []hdb200.Beam {
	b := make([]hdb200.Beam, len(m.Beams))
	for i, p := range m.Beams {
		b[i] = hdb200.Beam{N: p.N, A: p.A, J: p.J, E: p.E}
	}
	return b
}

func sections(m *Model) []float64 {
	s := make([]float64, 3*len(m.Beams))
	for i, p := range m.Beams {
		s[3*i], s[3*i+1], s[3*i+2] = p.A, p.J, p.E
	}
	return s
}

// prepare is model.go:336-355: Pins filled when empty (this MUTATES m, as the reference does), then the input checks.
func prepare(m *Model) error {
	if len(m.Pins) == 0 {
		m.Pins = make([][6]bool, len(m.Beams))
	}
	return hdb200.CheckModel(m.Points, beams(m), m.Pins, m.Supports)
}

func wrap(name string, err *error) { // model.go:372-385: error tree name + recover
	if r := recover(); r != nil {
		*err = fmt.Errorf("%s: stacktrace from panic: %v\n%s", name, r, debug.Stack())
		return
	}
	if *err != nil {
		*err = fmt.Errorf("%s: %w", name, *err)
	}
}

// LinearStatic is hd.LinearStatic (model.go:358-605).
func LinearStatic(out io.Writer, m *Model, lcs ...*LoadCase) (err error) {
	if out == nil {
		out = new(bytes.Buffer) // model.go:359-363: silent
	}
	for i := range lcs {
		lcs[i].reset()
	}
	const name = "Linear Elastic Analysis"
	fmt.Fprintf(out, "%s\n", name)
	defer wrap(name, &err)
	if err = prepare(m); err != nil {
		return fmt.Errorf("Prepared input data: %w", err)
	}
	nP, nB, dof := len(m.Points), len(m.Beams), 3*len(m.Points)
	for i := range lcs { // LoadCase.checkInputData, check.go:127-163
		ln := make([]hdb200.LoadNode, len(lcs[i].LoadNodes))
		for j, l := range lcs[i].LoadNodes {
			ln[j] = hdb200.LoadNode{N: l.N, Forces: l.Forces}
		}
		if e := hdb200.CheckLoads(nP, ln); e != nil {
			return fmt.Errorf("Prepared input data: %w", e)
		}
	}
	topo, err := hdb200.NewTopology(m.Points, beams(m), m.Pins, m.Supports) // was getK: assembly + pattern (model.go:318-333)
	if err != nil {
		return fmt.Errorf("Assembly node load: %w", err)
	}
	defer topo.Close()
	nLC := len(lcs)
	if nLC == 0 {
		return nil
	}
	loads := make([]float64, nLC*dof)
	for i, lc := range lcs { // assemblyNodeLoad, model.go:683-698
		for _, ln := range lc.LoadNodes {
			for k := 0; k < 3; k++ {
				loads[i*dof+ln.N*3+k] += ln.Forces[k]
			}
		}
	}
	disp, forces, react := make([]float64, nLC*dof), make([]float64, nLC*nB*6), make([]float64, nLC*dof)
	status, fail := []int32{0}, []int32{-1}
	sec := sections(m)
	if err = topo.StaticBatch(1, nLC, sec, loads, disp, forces, react, status, fail); err != nil {
		return err
	}
	upto := nLC
	switch status[0] {
	case statusOK:
	case statusIgnoreLoad: // model.go:424-437: the cases before the offending one keep their results
		upto, err = int(fail[0]), errIgnoreLoad
	case statusNonFinite:
		if fail[0] >= 0 { // assemblyNodeLoad of that case (model.go:692-696)
			upto, err = int(fail[0]), fmt.Errorf("Assembly node load: not valid node load")
		} else {
			upto, err = 0, fmt.Errorf("Assembly node load: not valid value")
		}
	case statusSingular:
		upto, err = 0, fmt.Errorf("Assembly node load: LU error factorization") // model.go:326-331
	default:
		upto, err = 0, fmt.Errorf("Assembly node load: status %d", status[0])
	}
	for i := 0; i < upto; i++ {
		lc := lcs[i]
		lc.PointDisplacementGlobal = make([][3]float64, nP)
		lc.Reactions = make([][3]float64, nP)
		lc.BeamForces = make([][6]float64, nB)
		for p := 0; p < nP; p++ {
			copy(lc.PointDisplacementGlobal[p][:], disp[i*dof+3*p:i*dof+3*p+3])
			copy(lc.Reactions[p][:], react[i*dof+3*p:i*dof+3*p+3])
		}
		for b := 0; b < nB; b++ {
			copy(lc.BeamForces[b][:], forces[(i*nB+b)*6:(i*nB+b)*6+6])
		}
		if amount := int(lc.AmountLinearBuckling); amount > 0 { // buckling block, model.go:491-601
			nOut, bst := []int32{0}, []int32{0}
			fac, shp := make([]float64, amount), make([]float64, amount*dof)
			if e := topo.BucklingBatch(1, 1, sec, forces[i*nB*6:(i+1)*nB*6], amount, amount, nOut, fac, shp, bst); e != nil {
				return e
			}
			if bst[0] != statusOK {
				return fmt.Errorf("Linear buckling: status %d", bst[0])
			}
			for r := 0; r < int(nOut[0]) && r < amount; r++ {
				br := BucklingResult{Factor: fac[r], PointDisplacementGlobal: make([][3]float64, nP)}
				for p := 0; p < nP; p++ {
					copy(br.PointDisplacementGlobal[p][:], shp[r*dof+3*p:r*dof+3*p+3])
				}
				lc.LinearBucklingResult = append(lc.LinearBucklingResult, br)
			}
		}
	}
	return err
}

// Modal is hd.Modal (model.go:719-851).
func Modal(out io.Writer, m *Model, mc *ModalCase) (err error) {
	if out == nil {
		out = new(bytes.Buffer) // model.go:720-724
	}
	mc.reset()
	const name = "Modal Analysis"
	fmt.Fprintf(out, "%s\n", name)
	defer wrap(name, &err)
	if err = prepare(m); err != nil {
		return err
	}
	nP, dof := len(m.Points), 3*len(m.Points)
	mm := make([]hdb200.ModalMass, len(mc.ModalMasses))
	massN, mass := make([]int, len(mm)), make([]float64, len(mm))
	for i, q := range mc.ModalMasses {
		mm[i] = hdb200.ModalMass{N: q.N, Mass: q.Mass}
		massN[i], mass[i] = q.N, q.Mass
	}
	if err = hdb200.CheckModal(nP, m.Supports, mm); err != nil { // check.go:411-452
		return err
	}
	if len(mm) == 0 {
		return nil // H = 0: no results and no error (model.go:772-805)
	}
	topo, err := hdb200.NewTopology(m.Points, beams(m), m.Pins, m.Supports)
	if err != nil {
		return fmt.Errorf("Assembly node load: %w", err)
	}
	defer topo.Close()
	s := topo.ModeCount(massN)
	if s <= 0 {
		return nil
	}
	nModes, status := []int32{0}, []int32{0}
	hz, shapes := make([]float64, s), make([]float64, s*dof)
	if err = topo.ModalBatch(1, sections(m), massN, mass, s, nModes, hz, shapes, status); err != nil {
		return err
	}
	if status[0] != statusOK {
		return fmt.Errorf("Eigen factorization is not ok") // model.go:810
	}
	for k := 0; k < int(nModes[0]); k++ {
		r := ModalResult{Hz: hz[k], ModalDisplacement: make([][3]float64, nP)}
		for p := 0; p < nP; p++ {
			copy(r.ModalDisplacement[p][:], shapes[k*dof+3*p:k*dof+3*p+3])
		}
		mc.Result = append(mc.Result, r)
	}
	sort.SliceStable(mc.Result, func(i, j int) bool { return mc.Result[i].Hz < mc.Result[j].Hz }) // model.go:846-848 (already ascending)
	return nil
}

// Run is hd.Run (model.go:926-969): LinearStatic over all cases, Modal per case, errors collected, not fatal.  The text
// report of the reference (Model.String, LoadCase.String, ModalCase.String) is written by the reference's own String
// methods, which the drop-in keeps; they are elided here.
func Run(out io.Writer, m *Model, lcs []LoadCase, mcs []ModalCase) error {
	var errs []error
	if len(lcs) > 0 {
		ptr := make([]*LoadCase, len(lcs))
		for i := range lcs {
			ptr[i] = &lcs[i]
		}
		if err := LinearStatic(out, m, ptr...); err != nil {
			errs = append(errs, err)
		}
	}
	for i := range mcs {
		if err := Modal(out, m, &mcs[i]); err != nil {
			errs = append(errs, err)
			continue
		}
	}
	if len(errs) > 0 {
		return fmt.Errorf("Run function: %w", errors.Join(errs...))
	}
	return nil
}

// BatchResult holds the flat result arrays of RunBatch: Disp/React [nModels][nLC][nP][3], Forces [nModels][nLC][nB][6],
// Status/FailCase per model (HDB_STATUS_*; one bad model does not poison the batch).
type BatchResult struct {
	Disp, Forces, React []float64
	Status, FailCase    []int32
}

// RunBatch solves nModels models that share m's topology: sections [nModels][nB][3] (A, J, E per beam) and
// loads [nModels][nLC][dof] (assemblyNodeLoad output).  New: the reference has no batched call.
func RunBatch(m *Model, nModels, nLC int, sectionsAll, loads []float64) (*BatchResult, error) {
	if err := prepare(m); err != nil {
		return nil, err
	}
	topo, err := hdb200.NewTopology(m.Points, beams(m), m.Pins, m.Supports)
	if err != nil {
		return nil, err
	}
	defer topo.Close()
	return runBatch(func(r *BatchResult) error {
		return topo.StaticBatch(nModels, nLC, sectionsAll, loads, r.Disp, r.Forces, r.React, r.Status, r.FailCase)
	}, m, nModels, nLC)
}

// RunBatchMulti is RunBatch over several GPUs of the box from this one process: contiguous model ranges per device,
// results written straight into the returned slices (hdb_multi_static_batch).
func RunBatchMulti(m *Model, devices []int32, nModels, nLC int, sectionsAll, loads []float64) (*BatchResult, error) {
	if err := prepare(m); err != nil {
		return nil, err
	}
	mt, err := hdb200.NewMulti(m.Points, beams(m), m.Pins, m.Supports, devices)
	if err != nil {
		return nil, err
	}
	defer mt.Close()
	return runBatch(func(r *BatchResult) error {
		return mt.StaticBatch(nModels, nLC, sectionsAll, loads, r.Disp, r.Forces, r.React, r.Status, r.FailCase)
	}, m, nModels, nLC)
}

func runBatch(call func(*BatchResult) error, m *Model, nModels, nLC int) (*BatchResult, error) {
	dof, nB := 3*len(m.Points), len(m.Beams)
	r := &BatchResult{
		Disp: make([]float64, nModels*nLC*dof), Forces: make([]float64, nModels*nLC*nB*6), React: make([]float64, nModels*nLC*dof),
		Status: make([]int32, nModels), FailCase: make([]int32, nModels),
	}
	if err := call(r); err != nil {
		return nil, err
	}
	return r, nil
}
