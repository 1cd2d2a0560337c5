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


def _rdzv_path():
    port = os.environ.get("MASTER_PORT", "0")
    run = os.environ.get("TORCHELASTIC_RUN_ID", "none")
    return "/tmp/bpk_rdzv_%s_%s.bin" % (port, run)


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
            uid = be.comm_unique_id()
            tmp = path + ".tmp.%d" % os.getpid()
            with open(tmp, "wb") as f:
                f.write(uid)
            os.replace(tmp, path)
        else:
            uid = None
            while time.time() - t_start < timeout:
                try:
                    st = os.stat(path)
                    if st.st_size == 128 and st.st_mtime > t_start - 120.0:
                        with open(path, "rb") as f:
                            uid = f.read()
                        break
                except FileNotFoundError:
                    pass
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
        tmp = "%s.ipc%d.tmp" % (path, r)
        with open(tmp, "wb") as f:
            f.write(mine)
        os.replace(tmp, "%s.ipc%d" % (path, r))
    except Exception:
        ok = 0.0
    handles = []
    t0 = time.time()
    for q in range(w):
        h = None
        while ok and time.time() - t0 < timeout:
            try:
                with open("%s.ipc%d" % (path, q), "rb") as f:
                    h = f.read()
                if len(h) == 64:
                    break
            except FileNotFoundError:
                pass
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
