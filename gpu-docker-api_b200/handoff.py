"""Python mirror of integration/go/services/vmig_handoff.go -- the "next" rows N1-N3 of SURVEY.md §8f.

The Go toolchain is absent in this image, so the hand-off logic a maintainer would add to
internal/services (snapshot per version, rollback from a snapshot, two-pass live hand-off) is kept here
line for line with the Go text, on top of the same C-ABI calls, so that the SEQUENCE is testable
(tests/test_gpu.py::test_handoff_*).  Names and layout follow the reference:

    merges/<rs>/<rs>-<v>/            setToMergeMap, internal/services/replicaset.go:681-704
    merges/<rs>/<rs>-<v>/diff        the snapshot of the version's UpperDir (the commented-out CopyDir there)
    merges/<rs>/<rs>-<v>/blocks.vmig its block table (new)

Docker lookups stay injected (set_resolver), pause/resume are callables (rs.PauseContainer,
internal/services/replicaset.go:641, and its inverse).
"""
from __future__ import annotations

import logging
import os
from pathlib import Path
from typing import Callable, Optional

from . import (F_PRUNE, F_VERIFY, GetContainerMergedLayer, GetVolumeMountPoint, VmigError, manifest, migrate_tree)

log = logging.getLogger("vmig.handoff")
_root: dict = {"cwd": None}


def set_merges_root(path) -> None:
    """The control plane uses its working directory (os.Getwd, replicaset.go:689); tests point it elsewhere."""
    _root["cwd"] = None if path is None else Path(path)


def versionDir(ctrVersionName: str) -> Path:
    """<cwd>/merges/<rs>/<rs>-<v> -- same rule as setToMergeMap (replicaset.go:689-692)."""
    cwd = _root["cwd"] or Path(os.getcwd())
    return cwd / "merges" / ctrVersionName.split("-")[0] / ctrVersionName


def snapshotPaths(ctrVersionName: str):
    d = versionDir(ctrVersionName)
    return d / "diff", d / "blocks.vmig"


def CopyDirDiff(src, dest, priorTable: Optional[os.PathLike], outTable: Optional[os.PathLike], flags: int = 0) -> dict:
    """utils.CopyDirDiff of the cgo shim: CopyDir with the block tables of the diff-skip path."""
    return migrate_tree(src, dest, priorTable or None, outTable or None, flags=flags)


def CopyDirDiffVerified(src, dest, priorTable, outTable) -> dict:
    """The pass after which the old container is deleted: the destination is re-read through the GPU and must hash
    like the source (VMIG_F_VERIFY), and holds nothing the source does not (VMIG_F_PRUNE)."""
    return CopyDirDiff(src, dest, priorTable, outTable, F_VERIFY | F_PRUNE)


def SnapshotVersion(ctrVersionName: str) -> dict:
    """N1: keep the layer of <rs>-<v> and the XXH64 table of its 4 MiB blocks before the container is deleted."""
    upper = GetContainerMergedLayer(ctrVersionName)
    data, table = snapshotPaths(ctrVersionName)
    data.mkdir(parents=True, exist_ok=True)
    return CopyDirDiff(upper, data, None, table)


def RollbackFromSnapshot(targetCtrVersionName: str, newContainer: str, seedTable=None) -> dict:
    """N1: fill newContainer's layer with version `target`.  seedTable = the table written FOR newContainer's layer by
    whatever put its present content there; only blocks that differ then travel.  A seed table that no longer
    describes the layer (failed earlier pass, out-of-band change) is detected per file by the engine and that file is
    copied in full; the result is verified either way."""
    data, _ = snapshotPaths(targetCtrVersionName)
    upper = GetContainerMergedLayer(newContainer)
    _, newTable = snapshotPaths(newContainer)
    newTable.parent.mkdir(parents=True, exist_ok=True)
    return CopyDirDiffVerified(data, upper, seedTable, newTable)


def HandoffCopy(oldContainer: str, newContainer: str, pause: Callable[[str], None], resume: Callable[[str], None]) -> dict:
    """N2: replaces utils.CopyOldMergedToNewContainerMerged in PatchContainer (replicaset.go:333).  The old container
    keeps running during pass 1 (the bulk of the bytes), is paused for pass 2, which re-reads the source and moves
    only the blocks whose hash changed in between, prunes what the tenant deleted meanwhile, verifies, and leaves the
    old container paused for the caller to delete.  Any engine error resumes the old container and is raised."""
    src = GetContainerMergedLayer(oldContainer)
    dst = GetContainerMergedLayer(newContainer)
    _, table = snapshotPaths(newContainer)
    table.parent.mkdir(parents=True, exist_ok=True)
    pass1: Optional[Path] = Path(str(table) + ".pass1")
    out = {"pass1": None, "pass1_error": None}
    try:
        try:
            out["pass1"] = CopyDirDiff(src, dst, None, pass1)
        except VmigError as e:      # live source: a file that shrinks under the reader, a directory swapped for a symlink
            log.warning("vmig: live pass of %s failed (%s); falling back to one paused pass", oldContainer, e)
            out["pass1_error"] = e.code
            pass1 = None
        pause(oldContainer)
        try:
            out["pass2"] = CopyDirDiffVerified(src, dst, pass1, table)
        except VmigError:
            resume(oldContainer)
            raise
        return out
    finally:
        if pass1 is not None and pass1.exists():
            pass1.unlink()


def UsedBytes(volVersionName: str) -> int:
    """N3: replaces utils.DirSize in PatchVolumeSize's shrink check (volume.go:126-140)."""
    return manifest(GetVolumeMountPoint(volVersionName))["bytes_total"]
