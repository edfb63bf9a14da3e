//go:build vmig

// vmig_handoff.go -- Go wiring for the "next" rows N1-N3 of SURVEY.md section 8f, on top of the cgo shim in
// utils/copy_vmig.go.  New file: nothing in it exists in the reference; the comments name the reference lines a
// maintainer changes to call it.  NOT COMPILED in this repository's image (no Go toolchain); the engine calls it
// makes are the ones tests/test_gpu.py exercises through the Python mirror gpu-docker-api_b200/handoff.py, which is
// kept line for line with this file (tests/test_gpu.py::test_handoff_*).
package services

/*
#cgo LDFLAGS: -lvmig -lstdc++ -lpthread -ldl -lrt
#include <stdlib.h>
#include <string.h>
#include <vmig.h>
*/
import "C"

import (
	"os"
	"path/filepath"
	"runtime"
	"strings"
	"unsafe"

	"github.com/ngaut/log"
	"github.com/pkg/errors"

	"github.com/mayooot/gpu-docker-api/utils"
)

// versionDir is where a ReplicaSet version keeps its artefacts; same rule as setToMergeMap
// (internal/services/replicaset.go:689-692): <cwd>/merges/<rs>/<rs>-<v>.  The snapshot of the diff layer goes
// to <dir>/diff, its block table to <dir>/blocks.vmig.
func versionDir(ctrVersionName string) string {
	cwd, _ := os.Getwd()
	return filepath.Join(cwd, "merges", strings.Split(ctrVersionName, "-")[0], ctrVersionName)
}

func snapshotPaths(ctrVersionName string) (data, table string) {
	d := versionDir(ctrVersionName)
	return filepath.Join(d, "diff"), filepath.Join(d, "blocks.vmig")
}

// N1 -- SnapshotVersion is the body the commented-out block of setToMergeMap (replicaset.go:684-701) was
// meant to have: keep the layer of <rs>-<v> and the XXH64 table of its 4 MiB blocks before the container is
// deleted.  Call it from setToMergeMap in place of the commented utils.CopyDir; ContainerMergeMap keeps
// pointing at versionDir.
func SnapshotVersion(ctrVersionName string) error {
	upper, err := utils.GetContainerMergedLayer(ctrVersionName)
	if err != nil {
		return errors.WithMessagef(err, "utils.GetContainerMergedLayer failed, container: %s", ctrVersionName)
	}
	data, table := snapshotPaths(ctrVersionName)
	if err := os.MkdirAll(data, 0755); err != nil {
		return errors.Wrapf(err, "mkdir %s", data)
	}
	return errors.WithMessagef(utils.CopyDirDiff(upper, data, "", table), "snapshot of %s failed", ctrVersionName)
}

// N1 -- RollbackFromSnapshot fills newContainer's layer with version `target` of the ReplicaSet.  The new
// layer was just created from the image, so it is seeded with a plain copy of the snapshot; when the caller
// has first put the CURRENT version there (seedTable = that version's blocks.vmig, e.g. after a failed
// roll-forward), only the blocks that differ between the two versions travel.  Replaces the CopyDir of
// RollbackContainer (replicaset.go:421).
func RollbackFromSnapshot(targetCtrVersionName, newContainer, seedTable string) error {
	data, _ := snapshotPaths(targetCtrVersionName)
	upper, err := utils.GetContainerMergedLayer(newContainer)
	if err != nil {
		return errors.WithMessagef(err, "utils.GetContainerMergedLayer failed, container: %s", newContainer)
	}
	_, newTable := snapshotPaths(newContainer)
	if err := os.MkdirAll(filepath.Dir(newTable), 0755); err != nil {
		return errors.Wrapf(err, "mkdir %s", filepath.Dir(newTable))
	}
	// verified + pruned: a seed table that no longer describes the layer is detected per file by the engine (inode/ctime
	// recorded in the table) and that file is copied in full; the re-hash of the destination is the backstop
	return errors.WithMessagef(utils.CopyDirDiffVerified(data, upper, seedTable, newTable), "rollback to %s failed", targetCtrVersionName)
}

// N2 -- HandoffCopy replaces utils.CopyOldMergedToNewContainerMerged in PatchContainer (replicaset.go:333):
// the old container keeps running during pass 1 (the bulk of the bytes), is paused for pass 2, which re-reads
// the source and moves only the blocks whose hash changed in between, and stays paused until the caller has
// started the new container and deleted the old one.  Any engine error resumes the old container and is
// returned BEFORE DeleteContainerForUpdate (replicaset.go:350) can run -- the reference ignores the tar pipe's
// exit status (utils/copy.go:23).  pause/resume are rs.PauseContainer (replicaset.go:641) and its inverse.
func HandoffCopy(oldContainer, newContainer string, pause, resume func(name string) error) error {
	src, err := utils.GetContainerMergedLayer(oldContainer)
	if err != nil {
		return errors.WithMessagef(err, "utils.GetContainerMergedLayer failed, container: %s", oldContainer)
	}
	dst, err := utils.GetContainerMergedLayer(newContainer)
	if err != nil {
		return errors.WithMessagef(err, "utils.GetContainerMergedLayer failed, container: %s", newContainer)
	}
	_, table := snapshotPaths(newContainer)
	if err := os.MkdirAll(filepath.Dir(table), 0755); err != nil {
		return errors.Wrapf(err, "mkdir %s", filepath.Dir(table))
	}
	pass1 := table + ".pass1"
	defer os.Remove(pass1)
	// pass 1, live: a file that shrinks under the reader fails with VMIG_ESRCCHANGED, a directory swapped for a
	// symlink is never followed (DESIGN.md, source-side path safety); both are retried once, paused.
	if err := utils.CopyDirDiff(src, dst, "", pass1); err != nil {
		log.Warnf("vmig: live pass of %s failed (%v); falling back to one paused pass", oldContainer, err)
		pass1 = ""
	}
	if err := pause(oldContainer); err != nil {
		return errors.WithMessage(err, "pause before the final pass failed")
	}
	if err := utils.CopyDirDiffVerified(src, dst, pass1, table); err != nil {
		_ = resume(oldContainer)
		return errors.WithMessage(err, "final pass failed; old container resumed, nothing deleted")
	}
	return nil
}

// N3 -- UsedBytes replaces utils.DirSize in PatchVolumeSize's shrink check (volume.go:126-140): the engine's
// metadata walk (no GPU work) returns the same sum of regular-file sizes; with a non-empty tablePath it also
// leaves the manifest the coming migration can reuse.
func UsedBytes(volVersionName string) (int64, error) {
	mountpoint, err := utils.GetVolumeMountPoint(volVersionName)
	if err != nil {
		return 0, errors.WithMessage(err, "utils.GetVolumeMountPoint failed")
	}
	cs := C.CString(mountpoint)
	defer C.free(unsafe.Pointer(cs))
	var st C.vmig_stats
	runtime.LockOSThread() // vmig_last_error() is thread-local: keep the call and the error fetch on one OS thread
	defer runtime.UnlockOSThread()
	if rc := C.vmig_manifest(cs, 0, 0, nil, &st); rc != 0 {
		return 0, errors.Errorf("vmig_manifest(%s): %s: %s", mountpoint, C.GoString(C.vmig_strerror(rc)), C.GoString(C.vmig_last_error()))
	}
	return int64(st.bytes_total), nil
}
