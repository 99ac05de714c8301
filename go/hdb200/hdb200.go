// Package hdb200 binds libhdb200.so (include/hdb200.h) for the hd package.
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no Go toolchain here (see DESIGN.md).  This is the
// binding a maintainer of Konstantin8105/hd adds; it is written against the same header the C++ and
// Python hosts use.  The reference keeps check.go, the text reports and its exported structs unchanged;
// only the bodies of LinearStatic (model.go:358-605) and Modal (model.go:719-851) change: prepare/check in
// Go -> flatten -> ONE cgo call per phase -> results copied into the Go result slices.
package hdb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../hd_b200 -lhdb200 -Wl,-rpath,${SRCDIR}/../../hd_b200
#include <stdlib.h>
#include "hdb200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"
)

// Beam mirrors hd.BeamProp (model.go:58-77): identical memory layout, so []hd.BeamProp can be passed
// to C without copying (all fields are pointer-free).
type Beam struct {
	N       [2]int
	A, J, E float64
}

// Topology is the handle of the symbolic phase (hdb_topology_create).
type Topology struct{ h *C.hdb_topology }

func errOf(rc C.int, where string) error {
	if rc >= 0 {
		return nil
	}
	return fmt.Errorf("%s: %s (%s)", where, C.GoString(C.hdb_error_string(rc)), C.GoString(C.hdb_last_error_detail()))
}

func b2u(b [][6]bool) *C.uint8_t {
	if len(b) == 0 {
		return nil // all rigid, model.go:344-346
	}
	return (*C.uint8_t)(unsafe.Pointer(&b[0][0])) // Go bool is one byte, 0 or 1
}

// NewTopology replaces the structural half of getK / assemblyK (model.go:318-333, 607-681).
func NewTopology(points [][2]float64, beams []Beam, pins [][6]bool, supports [][3]bool) (*Topology, error) {
	t := &Topology{}
	rc := C.hdb_topology_create(C.int32_t(len(points)), (*C.double)(unsafe.Pointer(&points[0][0])),
		C.int32_t(len(beams)), (*C.hdb_beam)(unsafe.Pointer(&beams[0])), b2u(pins),
		(*C.uint8_t)(unsafe.Pointer(&supports[0][0])), nil, &t.h)
	if err := errOf(rc, "hdb_topology_create"); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(t, func(t *Topology) { t.Close() })
	return t, nil
}

// Close releases the handle (explicit lifetime; the finalizer is only a safety net).
func (t *Topology) Close() {
	if t.h != nil {
		C.hdb_topology_destroy(t.h)
		t.h = nil
	}
}

// StaticBatch is LinearStatic for nModels models sharing the topology (model.go:401-490).
// sections [nModels][nB][3], loads [nModels][nLC][dof] (assemblyNodeLoad output); the result slices are
// written in place: disp/react [nModels][nLC][nP][3], forces [nModels][nLC][nB][6].
func (t *Topology) StaticBatch(nModels, nLC int, sections, loads, disp, forces, react []float64,
	status, failCase []int32) error {
	p := func(s []float64) *C.double {
		if len(s) == 0 {
			return nil
		}
		return (*C.double)(unsafe.Pointer(&s[0]))
	}
	rc := C.hdb_static_batch(t.h, C.int32_t(nModels), p(sections), C.int32_t(nLC), p(loads), p(disp), p(forces),
		p(react), (*C.int32_t)(unsafe.Pointer(&status[0])), (*C.int32_t)(unsafe.Pointer(&failCase[0])))
	runtime.KeepAlive(sections)
	runtime.KeepAlive(loads)
	return errOf(rc, "hdb_static_batch")
}

// StaticBatchGeom is StaticBatch with per-model coordinates points[nModels][nP][2] (hdb_static_batch_geom): the
// geometry checks of check.go (NaN / Inf coordinates, beams of length zero) come back per model in status.
func (t *Topology) StaticBatchGeom(nModels, nLC int, points, sections, loads, disp, forces, react []float64,
	status, failCase []int32) error {
	p := func(s []float64) *C.double {
		if len(s) == 0 {
			return nil
		}
		return (*C.double)(unsafe.Pointer(&s[0]))
	}
	rc := C.hdb_static_batch_geom(t.h, C.int32_t(nModels), p(points), p(sections), C.int32_t(nLC), p(loads), p(disp),
		p(forces), p(react), (*C.int32_t)(unsafe.Pointer(&status[0])), (*C.int32_t)(unsafe.Pointer(&failCase[0])))
	runtime.KeepAlive(points)
	runtime.KeepAlive(sections)
	runtime.KeepAlive(loads)
	return errOf(rc, "hdb_static_batch_geom")
}

// ModalBatch is Modal for nModels models (model.go:719-851). massN is shared, mass [nModels][len(massN)].
func (t *Topology) ModalBatch(nModels int, sections []float64, massN []int, mass []float64, maxModes int,
	nModes []int32, hz, shapes []float64, status []int32) error {
	rc := C.hdb_modal_batch(t.h, C.int32_t(nModels), (*C.double)(unsafe.Pointer(&sections[0])), C.int32_t(len(massN)),
		(*C.int64_t)(unsafe.Pointer(&massN[0])), (*C.double)(unsafe.Pointer(&mass[0])), C.int32_t(maxModes),
		(*C.int32_t)(unsafe.Pointer(&nModes[0])), (*C.double)(unsafe.Pointer(&hz[0])),
		(*C.double)(unsafe.Pointer(&shapes[0])), (*C.int32_t)(unsafe.Pointer(&status[0])))
	return errOf(rc, "hdb_modal_batch")
}

// ModeCount = number of distinct free mass DOFs = modes Modal returns.
func (t *Topology) ModeCount(massN []int) int {
	return int(C.hdb_modal_mode_count(t.h, C.int32_t(len(massN)), (*C.int64_t)(unsafe.Pointer(&massN[0]))))
}

// BucklingBatch is the buckling block of LinearStatic (model.go:491-601) for nModels models and nLC load cases:
// forces [nModels][nLC][nB][6] are the beam end forces StaticBatch returned; per (model, case) nOut factors
// (ascending, at most amount) and their shapes [maxOut][nP][3].
func (t *Topology) BucklingBatch(nModels, nLC int, sections, forces []float64, amount, maxOut int,
	nOut []int32, factors, shapes []float64, status []int32) error {
	p := func(s []float64) *C.double {
		if len(s) == 0 {
			return nil
		}
		return (*C.double)(unsafe.Pointer(&s[0]))
	}
	rc := C.hdb_buckling_batch(t.h, C.int32_t(nModels), p(sections), C.int32_t(nLC), p(forces), C.int32_t(amount),
		C.int32_t(maxOut), (*C.int32_t)(unsafe.Pointer(&nOut[0])), p(factors), p(shapes),
		(*C.int32_t)(unsafe.Pointer(&status[0])))
	return errOf(rc, "hdb_buckling_batch")
}

// UniformLoadsDev builds loads[nModels][nLC][dof] on the device from mod.LoadUniform parameters (mod/uniform.go:46-211):
// q and f are device pointers ([nModels][nLC][len(beams)][2], [nModels][nLC][len(nodes)][3]), loads a device pointer.
func (t *Topology) UniformLoadsDev(nModels, nLC int, beams []int32, projection []bool, q unsafe.Pointer, nodes []int64,
	f unsafe.Pointer, loads unsafe.Pointer) error {
	pr := make([]C.uint8_t, len(projection))
	for i, b := range projection {
		if b {
			pr[i] = 1
		}
	}
	var pb *C.int32_t
	var pp *C.uint8_t
	var pn *C.int64_t
	if len(beams) > 0 {
		pb, pp = (*C.int32_t)(unsafe.Pointer(&beams[0])), &pr[0]
	}
	if len(nodes) > 0 {
		pn = (*C.int64_t)(unsafe.Pointer(&nodes[0]))
	}
	rc := C.hdb_uniform_loads_dev(t.h, C.int32_t(nModels), C.int32_t(nLC), C.int32_t(len(beams)), pb, pp, (*C.double)(q),
		C.int32_t(len(nodes)), pn, (*C.double)(f), (*C.double)(loads), nil)
	return errOf(rc, "hdb_uniform_loads_dev")
}

// CheckModel is checkInputData of check.go:12-109, 223-397 restated in the library (hosts that keep the reference's
// own check.go do not need it).  A non-nil error carries the library's list of findings.
func CheckModel(points [][2]float64, beams []Beam, pins [][6]bool, supports [][3]bool) error {
	msg := make([]byte, 4096)
	var pp *C.double
	var pb *C.hdb_beam
	var ps *C.uint8_t
	if len(points) > 0 {
		pp = (*C.double)(unsafe.Pointer(&points[0][0]))
	}
	if len(beams) > 0 {
		pb = (*C.hdb_beam)(unsafe.Pointer(&beams[0]))
	}
	if len(supports) > 0 {
		ps = (*C.uint8_t)(unsafe.Pointer(&supports[0][0]))
	}
	rc := C.hdb_check_model(C.int32_t(len(points)), pp, C.int32_t(len(beams)), pb, C.int32_t(len(pins)), b2u(pins),
		C.int32_t(len(supports)), ps, (*C.char)(unsafe.Pointer(&msg[0])), C.int32_t(len(msg)))
	if rc == 0 {
		return nil
	}
	return fmt.Errorf("%s", C.GoString((*C.char)(unsafe.Pointer(&msg[0]))))
}

// LoadNode mirrors hd.LoadNode (model.go:304-315), 32 bytes, pointer-free.
type LoadNode struct {
	N      int
	Forces [3]float64
}

// ModalMass mirrors hd.ModalMass (model.go:294-301), 16 bytes, pointer-free.
type ModalMass struct {
	N    int
	Mass float64
}

// CheckLoads is LoadCase.checkInputData (check.go:127-163).
func CheckLoads(nP int, loads []LoadNode) error {
	msg := make([]byte, 2048)
	var p *C.hdb_load_node
	if len(loads) > 0 {
		p = (*C.hdb_load_node)(unsafe.Pointer(&loads[0]))
	}
	if rc := C.hdb_check_loads(C.int32_t(nP), C.int32_t(len(loads)), p, (*C.char)(unsafe.Pointer(&msg[0])), C.int32_t(len(msg))); rc != 0 {
		return fmt.Errorf("%s", C.GoString((*C.char)(unsafe.Pointer(&msg[0]))))
	}
	return nil
}

// CheckModal is ModalCase.checkInputData (check.go:411-452).
func CheckModal(nP int, supports [][3]bool, masses []ModalMass) error {
	msg := make([]byte, 2048)
	var p *C.hdb_modal_mass
	var ps *C.uint8_t
	if len(masses) > 0 {
		p = (*C.hdb_modal_mass)(unsafe.Pointer(&masses[0]))
	}
	if len(supports) > 0 {
		ps = (*C.uint8_t)(unsafe.Pointer(&supports[0][0]))
	}
	if rc := C.hdb_check_modal(C.int32_t(nP), C.int32_t(len(supports)), ps, C.int32_t(len(masses)), p,
		(*C.char)(unsafe.Pointer(&msg[0])), C.int32_t(len(msg))); rc != 0 {
		return fmt.Errorf("%s", C.GoString((*C.char)(unsafe.Pointer(&msg[0]))))
	}
	return nil
}

// Multi drives several GPUs of one box from this one process (hdb_multi_*): contiguous model ranges per device.
type Multi struct{ h *C.hdb_multi }

// NewMulti runs the symbolic phase once per device; devices are CUDA ordinals.
func NewMulti(points [][2]float64, beams []Beam, pins [][6]bool, supports [][3]bool, devices []int32) (*Multi, error) {
	m := &Multi{}
	rc := C.hdb_multi_create(C.int32_t(len(points)), (*C.double)(unsafe.Pointer(&points[0][0])), C.int32_t(len(beams)),
		(*C.hdb_beam)(unsafe.Pointer(&beams[0])), b2u(pins), (*C.uint8_t)(unsafe.Pointer(&supports[0][0])), nil,
		C.int32_t(len(devices)), (*C.int32_t)(unsafe.Pointer(&devices[0])), &m.h)
	if err := errOf(rc, "hdb_multi_create"); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(m, func(m *Multi) { m.Close() })
	return m, nil
}

// Close releases every per-device handle.
func (m *Multi) Close() {
	if m.h != nil {
		C.hdb_multi_destroy(m.h)
		m.h = nil
	}
}

// StaticBatch is Topology.StaticBatch over all devices: host slices in, host slices out, one host thread per device
// inside the library, results bit-identical to one device.
func (m *Multi) StaticBatch(nModels, nLC int, sections, loads, disp, forces, react []float64, status, failCase []int32) error {
	p := func(s []float64) *C.double {
		if len(s) == 0 {
			return nil
		}
		return (*C.double)(unsafe.Pointer(&s[0]))
	}
	rc := C.hdb_multi_static_batch(m.h, C.int32_t(nModels), p(sections), C.int32_t(nLC), p(loads), p(disp), p(forces), p(react),
		(*C.int32_t)(unsafe.Pointer(&status[0])), (*C.int32_t)(unsafe.Pointer(&failCase[0])))
	runtime.KeepAlive(sections)
	runtime.KeepAlive(loads)
	return errOf(rc, "hdb_multi_static_batch")
}
