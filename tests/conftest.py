import os
import shutil
import sys
import tempfile
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def vm():
    """The package under test (gpu-docker-api_b200), built if necessary."""
    import __graft_entry__ as g
    if not (g.PKG_DIR / "libvmig.so").exists():
        g.build()
    return g.load_pkg()


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture()
def shm_tmp():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = Path(tempfile.mkdtemp(prefix="vmig_t_", dir=base))
    yield d
    shutil.rmtree(d, ignore_errors=True)


def make_rich_tree(root: Path, orc, seed: int = 5, big: int = (9 << 20) + 777):
    """A small tree with every entry kind tar reproduces."""
    (root / "sub" / "deep").mkdir(parents=True)
    (root / ".hidden_dir").mkdir()
    (root / "emptydir").mkdir()
    (root / "big.bin").write_bytes(orc.splitmix_bytes(seed, big).tobytes())
    (root / "exact.bin").write_bytes(orc.splitmix_bytes(seed + 1, 4 << 20).tobytes())
    (root / "sub" / "small.txt").write_bytes(b"x" * 4097)
    (root / "sub" / "deep" / "tiny").write_bytes(b"abc")
    (root / ".dotfile").write_bytes(b"dot")
    (root / ".hidden_dir" / "inner").write_bytes(orc.splitmix_bytes(seed + 2, 70000).tobytes())
    (root / "empty").write_bytes(b"")
    (root / "zeros.bin").write_bytes(bytes(5 << 20))
    os.symlink("big.bin", root / "lnk")
    os.symlink("/nonexistent/target", root / "sub" / "dangling")
    os.link(root / "big.bin", root / "sub" / "hard1")
    os.link(root / "sub" / "small.txt", root / "a_hard2")
    os.mkfifo(root / "fifo")
    if os.geteuid() == 0:
        try:
            os.mknod(root / "whiteout", 0o020000 | 0o644, os.makedev(0, 0))   # overlay2 whiteout
        except PermissionError:       # an overlayfs mount refuses to create its own whiteout marker by hand
            os.mknod(root / "whiteout", 0o020000 | 0o644, os.makedev(1, 3))
        os.chown(root / "sub", 1234, 4321)
        os.chown(root / "sub" / "small.txt", 1000, 1000)
        os.lchown(root / "lnk", 42, 43)
    os.chmod(root / "big.bin", 0o4750)
    os.chmod(root / "sub" / "deep", 0o711)
    os.chmod(root / "emptydir", 0o1777)
    t = 1_577_934_245_123_456_789   # 2020-01-02T03:04:05.123456789Z
    for p in [root / "big.bin", root / "sub" / "small.txt", root / "sub", root / "emptydir", root / "empty"]:
        os.utime(p, ns=(t, t))
    os.utime(root / "lnk", ns=(t + 10**9, t + 10**9), follow_symlinks=False)
    os.utime(root, ns=(t - 10**9, t - 10**9 + 5))
