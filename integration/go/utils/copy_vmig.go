//go:build vmig

package utils

/*
#cgo LDFLAGS: -lvmig -lstdc++ -lpthread -ldl -lrt
#include <stdlib.h>
#include <vmig.h>
*/
import "C"

import (
	"runtime"
	"unsafe"

	"github.com/ngaut/log"
	"github.com/pkg/errors"
)

// vmigErr turns a return code into an error.  vmig_last_error() is thread-local and is read by a SEPARATE cgo call:
// between two cgo calls the Go scheduler may move the goroutine to another OS thread, so every caller below brackets
// "call + vmigErr" with runtime.LockOSThread()/UnlockOSThread() (lockedCall).
func vmigErr(rc C.int, what string) error {
	if rc == 0 {
		return nil
	}
	return errors.Errorf("%s: %s (%d): %s", what, C.GoString(C.vmig_strerror(rc)), int(rc),
		C.GoString(C.vmig_last_error()))
}

// lockedCall runs f (one libvmig call) and fetches its error text on the same OS thread.
func lockedCall(what string, f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	return vmigErr(f(), what)
}

// CopyDir replaces `sh -c "(cd src; tar c .) | (cd dest; tar x)"` (utils/copy.go:17-27).
func CopyDir(src, dest string) error {
	cs, cd := C.CString(src), C.CString(dest)
	defer C.free(unsafe.Pointer(cs))
	defer C.free(unsafe.Pointer(cd))
	return errors.Wrapf(lockedCall("vmig_copy_dir", func() C.int { return C.vmig_copy_dir(cs, cd) }),
		"vmig copy failed, src:%s, dest: %s", src, dest)
}

// CopyDirDiff is CopyDir with the block tables the diff-skip path needs. prior/out are paths
// under the per-version directory setToMergeMap already creates
// (internal/services/replicaset.go:681-704): merges/<rs>/<rs>-<v>/blocks.vmig.
func CopyDirDiff(src, dest, priorTable, outTable string) error {
	return copyDirDiff(src, dest, priorTable, outTable, 0)
}

// CopyDirDiffVerified additionally re-reads the destination through the GPU and compares block tables
// (VMIG_F_VERIFY), and removes every destination entry the source no longer has (VMIG_F_PRUNE: a file the tenant
// deleted or renamed since an earlier pass must not reappear), before it reports success: for the pass after which
// the old container is deleted, and for rollbacks onto a seeded layer.
func CopyDirDiffVerified(src, dest, priorTable, outTable string) error {
	return copyDirDiff(src, dest, priorTable, outTable, C.VMIG_F_VERIFY|C.VMIG_F_PRUNE)
}

func copyDirDiff(src, dest, priorTable, outTable string, flags C.uint32_t) error {
	cs, cd := C.CString(src), C.CString(dest)
	defer C.free(unsafe.Pointer(cs))
	defer C.free(unsafe.Pointer(cd))
	var cp, co *C.char
	if priorTable != "" {
		cp = C.CString(priorTable)
		defer C.free(unsafe.Pointer(cp))
	}
	if outTable != "" {
		co = C.CString(outTable)
		defer C.free(unsafe.Pointer(co))
	}
	var st C.vmig_stats
	var o C.vmig_opts // zero value = defaults (all GPUs, 4 MiB blocks)
	o.flags = flags
	err := lockedCall("vmig_migrate_tree", func() C.int { return C.vmig_migrate_tree(cs, cd, cp, co, &o, &st) })
	if err == nil {
		log.Infof("vmig: %d bytes, %d/%d blocks skipped, %d files copied in full (stale table), %d pruned, %.2f GiB/s",
			uint64(st.bytes_total), uint64(st.blocks_skipped), uint64(st.blocks_total), uint64(st.files_untrusted),
			uint64(st.pruned), float64(st.bytes_total)/float64(st.ns_total)*1e9/(1<<30))
	}
	return err
}

// CopyOldMergedToNewContainerMerged (utils/copy.go:31-46): unchanged except for the callee.
// GetContainerMergedLayer (Docker ContainerInspect -> UpperDir, utils/copy.go:48-54) stays in Go.

// CopyOldMountPointToContainerMountPoint replaces moveVolumeData's helper container + `mv`
// (utils/copy.go:74-128): resolve both volume names with GetVolumeMountPoint (utils/copy.go:65-72,
// kept) and move on the host paths, synchronously, with the status checked.
func CopyOldMountPointToContainerMountPoint(oldVolume, newVolume string) error {
	src, err := GetVolumeMountPoint(oldVolume)
	if err != nil {
		return errors.WithMessage(err, "GetVolumeMountPoint failed")
	}
	dst, err := GetVolumeMountPoint(newVolume)
	if err != nil {
		return errors.WithMessage(err, "GetVolumeMountPoint failed")
	}
	cs, cd := C.CString(src), C.CString(dst)
	defer C.free(unsafe.Pointer(cs))
	defer C.free(unsafe.Pointer(cd))
	return errors.WithMessage(lockedCall("vmig_move_dir", func() C.int { return C.vmig_move_dir(cs, cd) }), "moveData failed")
}
