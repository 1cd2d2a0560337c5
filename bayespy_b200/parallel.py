"""Multi-GPU plumbing: one process per GPU, the sample-plate axis block-sharded over
ranks, one NCCL all-reduce of the plate-summed sufficient statistics per sweep
(SURVEY.md §8e).  The reference has no distributed path; the reduction site is the
plate-sum of Node._message_to_parent (node.py:650) / the einsum of dot.py:581.

No torch here: the NCCL unique id travels through a file in /tmp (all ranks are on
one node, as launched by ``python -m torch.distributed.run --nnodes=1``), and
barriers / max-over-ranks reductions are done with the library's own all-reduce.
"""
import os
import time

import numpy as np

from . import _bpk
from .darray import DArray

_state = dict(world=1, rank=0, ready=False)


def world():
    return _state["world"]


def rank():
    return _state["rank"]


def shard_bounds(n_total, world_size, r):
    """Contiguous block [n0, n1) of the plate axis owned by rank r (sizes differ by <= 1)."""
    base, rem = divmod(int(n_total), int(world_size))
    n0 = r * base + min(r, rem)
    return n0, n0 + base + (1 if r < rem else 0)


def _rdzv_dir():
    """Per-user 0700 directory for the rendezvous files (NCCL id, IPC handles)."""
    d = os.path.join(os.environ.get("BPK_RDZV_DIR", "/tmp"), "bpk_rdzv_%d" % os.getuid())
    try:
        os.mkdir(d, 0o700)
    except FileExistsError:
        pass
    st = os.lstat(d)
    import stat
    if not stat.S_ISDIR(st.st_mode) or st.st_uid != os.getuid() or (st.st_mode & 0o077):
        raise RuntimeError("rendezvous directory %s is not a private directory of this user" % d)
    return d


def _rdzv_path():
    """One file name per LAUNCH: every rank of one torchrun launch shares the agent's pid (our parent), the
    port and the restart count, so a file left behind by a crashed earlier launch can never be picked up."""
    port = os.environ.get("MASTER_PORT", "0")
    run = os.environ.get("TORCHELASTIC_RUN_ID", "none")
    restart = os.environ.get("TORCHELASTIC_RESTART_COUNT", "0")
    launch = os.environ.get("BPK_LAUNCH_ID", str(os.getppid()))
    safe = "".join(c if c.isalnum() else "_" for c in "%s_%s_%s_%s" % (port, run, restart, launch))
    return os.path.join(_rdzv_dir(), "id_%s.bin" % safe)


def _write_private(path, payload):
    """Create ``path`` atomically with mode 0600 (O_EXCL on a temporary name, then rename)."""
    tmp = "%s.tmp.%d" % (path, os.getpid())
    try:
        os.unlink(tmp)
    except FileNotFoundError:
        pass
    fd = os.open(tmp, os.O_WRONLY | os.O_CREAT | os.O_EXCL, 0o600)
    try:
        os.write(fd, payload)
    finally:
        os.close(fd)
    os.replace(tmp, path)


def _read_private(path, size):
    """Contents of ``path`` if it is a regular file of this user with exactly ``size`` bytes, else None."""
    try:
        fd = os.open(path, os.O_RDONLY | os.O_NOFOLLOW)
    except (FileNotFoundError, OSError):
        return None
    try:
        st = os.fstat(fd)
        import stat
        if not stat.S_ISREG(st.st_mode) or st.st_uid != os.getuid() or st.st_size != size:
            return None
        return os.read(fd, size)
    finally:
        os.close(fd)


def init_from_env(timeout=300.0):
    """Bind this process to cuda:LOCAL_RANK and create the NCCL communicator described by
    RANK / WORLD_SIZE (torchrun's environment).  Single-process runs are a no-op."""
    if _state["ready"]:
        return _state["world"], _state["rank"]
    w = int(os.environ.get("WORLD_SIZE", "1"))
    r = int(os.environ.get("RANK", "0"))
    be = _bpk.get()              # device = LOCAL_RANK (see _bpk.get)
    if w > 1:
        # NCCL writes its version / debug lines to stdout by default; keep stdout for the caller's own output
        # (bench.py prints exactly one JSON line there)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        path = _rdzv_path()
        t_start = time.time()
        if r == 0:
            try:
                os.unlink(path)                       # never let a peer read an older launch's id
            except FileNotFoundError:
                pass
            uid = be.comm_unique_id()
            _write_private(path, uid)
        else:
            uid = None
            while time.time() - t_start < timeout:
                uid = _read_private(path, 128)
                if uid is not None:
                    break
                time.sleep(0.05)
            if uid is None:
                raise RuntimeError("rank %d: no NCCL id at %s after %.0f s" % (r, path, timeout))
        be.comm_init(uid, w, r)
        _state.update(world=w, rank=r, ready=True)
        barrier()
        _open_peer_windows(be, w, r, path, timeout)
        if r == 0:
            try:
                os.remove(path)
            except OSError:
                pass
    _state.update(world=w, rank=r, ready=True)
    return w, r


def _open_peer_windows(be, w, r, path, timeout):
    """Exchange CUDA IPC handles of the peer-memory windows (files next to the NCCL id) so that the
    resident loop can all-reduce its statistics inside the sweep kernel over NVLink.  Falls back to
    the NCCL all-reduce (with a warning) if IPC is not available; BPK_NO_P2P=1 forces the fallback."""
    _state["p2p"] = False
    if os.environ.get("BPK_NO_P2P") or w > 8 or not hasattr(be, "xchg_create"):
        return
    ok = 1.0
    try:
        mine = be.xchg_create()
        _write_private("%s.ipc%d" % (path, r), mine)
    except Exception:
        ok = 0.0
    handles = []
    t0 = time.time()
    for q in range(w):
        h = None
        while ok and time.time() - t0 < timeout:
            h = _read_private("%s.ipc%d" % (path, q), 64)
            if h is not None:
                break
            time.sleep(0.02)
        if h is None or len(h) != 64:
            ok = 0.0
            h = b"\0" * 64
        handles.append(h)
    if ok:
        try:
            be.xchg_open(handles, w, r)
        except Exception as e:
            ok = 0.0
            import sys
            sys.stderr.write("bayespy_b200: peer-memory windows unavailable (%s); using NCCL all-reduce\n" % e)
    allok = float(np.min(allgather_scalar(ok)))
    barrier()
    try:
        os.remove("%s.ipc%d" % (path, r))
    except OSError:
        pass
    if allok < 1.0:
        if ok:
            be.xchg_close()
        return
    _state["p2p"] = True


def bind_to_gpu_numa():
    """Pin this process to the CPUs that are local to its GPU (``/sys/bus/pci/devices/<id>/local_cpulist``), so
    that pinned staging buffers are first-touched on the GPU's NUMA node and H2D copies do not cross the socket
    interconnect (an 8-GPU box has 4 GPUs per socket; an unpinned rank lost ~30 % of its copy bandwidth).
    Returns the CPU set, or None when the topology cannot be read."""
    be = _bpk.get()
    if not hasattr(be, "pci_bus_id") or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        bus = be.pci_bus_id()
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bus) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def set_world_for_testing(world_size, r):
    """TEST INFRASTRUCTURE: declare the sharding without NCCL (the oracle backend's
    all-reduce hook provides the collective, e.g. gloo on CPU)."""
    _state.update(world=world_size, rank=r, ready=True)


def allreduce_sum(arr):
    """In-place sum over ranks of a contiguous fp64 device array (identity when world == 1)."""
    if _state["world"] > 1:
        _bpk.get().allreduce_sum_f64(arr.ptr, arr.size)
    return arr


def allgather_scalar(x):
    """Every rank's value of a host float, via a one-hot all-reduce."""
    w, r = _state["world"], _state["rank"]
    if w == 1:
        return np.array([float(x)])
    v = np.zeros(w)
    v[r] = float(x)
    d = DArray.from_numpy(v)
    allreduce_sum(d)
    return d.numpy()


def barrier():
    if _state["world"] > 1:
        allgather_scalar(0.0)
    _bpk.get().sync()


def barrier_aligned(lead_s=0.002):
    """barrier(), then every rank leaves at the same instant.

    A collective releases its ranks tens to hundreds of microseconds apart (each host wakes from its own stream
    synchronisation).  For a timed region of a few milliseconds that skew is measured as work: the span of the rank
    that left first covers the wait for the rank that left last.  The ranks of one node share CLOCK_MONOTONIC, so they
    agree on a deadline (latest clock reading + lead_s) and spin up to it.  Single node only (as torchrun --nnodes=1)."""
    barrier()
    if _state["world"] > 1:
        import time
        deadline = float(np.max(allgather_scalar(time.monotonic()))) + lead_s
        while time.monotonic() < deadline:
            pass
