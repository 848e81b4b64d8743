"""CollisionEngine: the batched, device-resident collision query behind the drop-in classes.

Thin Python over the C-ABI (include/jogramop_b200.h); torch is used only for device memory and the
current CUDA stream.  All batched methods are pure functions of their inputs.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

SCENE, PLANE, SELF, LIMITS, NO_EPA = 1, 2, 4, 8, 16
POSE_MAT34, POSE_POS_QUAT = 0, 1


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def joint_qidx(robot):
    """Joint -> configuration-vector slot (arm_joint_ids order, simulation.py:159); fingers -> -2."""
    qidx = np.full(robot.n_links, -1, np.int32)
    for k, j in enumerate(robot.arm_joint_ids):
        qidx[j] = k
    for j in robot.finger_joint_ids:
        qidx[j] = -2
    return qidx


def unpack_bits(words, n):
    """Packed verdict words (bit i%32 of word i//32) -> bool tensor [n] on the same device."""
    w = words.view(torch.int32)
    shifts = torch.arange(32, device=w.device, dtype=torch.int32)
    return ((w[:, None] >> shifts[None, :]) & 1).reshape(-1)[:n].to(torch.bool)


class CollisionEngine:
    def __init__(self, robot, scene=None, device=0, threshold=-0.001, query_distance=0.01, finger_open=0.04,
                 chunk=0, mesh_mode=False, mesh_margin=0.001, lanes=1):
        """mesh_mode: check against the scene's merged triangle mesh (``scene.tri_verts`` / ``scene.tris`` = the
        reference's export/obstacles.obj) through a GPU-built LBVH instead of against the convex pieces.
        lanes: scratch arenas a call of more than one pass may use side by side (``jg_engine_set_lanes``)."""
        L = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.JogramopError('no CUDA device: the jogramop_framework_b200 engine has no CPU fallback')
        self.robot, self.scene = robot, scene
        self.dof = len(robot.arm_joint_ids)
        self.device = torch.device('cuda', device if isinstance(device, int) else torch.device(device).index or 0)
        self._L = L
        lim = _f64(robot.joint_limits[robot.arm_joint_ids])
        k = [_i32(robot.joint_type), _i32(robot.joint_parent), _f64(robot.joint_origin[:, :3, :]),
             _f64(robot.joint_axis), joint_qidx(robot), lim, _i32(robot.shape_link), _f64(robot.shape_margin),
             _i32(robot.shape_vert_offset), _f64(robot.core_verts), _i32(robot.shape_plane_offset),
             _f64(robot.plane_verts), _f64(robot.shape_plane_margin), _i32(robot.self_pairs)]
        self._r = L.jg_robot_create(robot.n_links, _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), _p(k[5]), self.dof,
                                    robot.n_shapes, _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]), _p(k[10]), _p(k[11]),
                                    _p(k[12]), len(robot.self_pairs), _p(k[13]))
        if not self._r:
            raise _lib.JogramopError('jg_robot_create: ' + L.jg_last_error().decode())
        self._s = None
        self.mesh_mode = bool(mesh_mode)
        if scene is not None and mesh_mode:
            ks = [_f64(scene.tri_verts), _i32(scene.tris), _f64(scene.plane)]
            self._s = L.jg_scene_create_mesh(len(ks[0]), _p(ks[0]), len(ks[1]), _p(ks[1]), float(mesh_margin),
                                             _p(ks[2]) if scene.has_plane else None)
            if not self._s:
                raise _lib.JogramopError('jg_scene_create_mesh: ' + L.jg_last_error().decode())
        elif scene is not None:
            ks = [_i32(scene.piece_vert_offset), _f64(scene.piece_verts), _f64(scene.piece_margin), _f64(scene.plane)]
            self._s = L.jg_scene_create(scene.n_pieces, _p(ks[0]), _p(ks[1]), _p(ks[2]),
                                        _p(ks[3]) if scene.has_plane else None)
            if not self._s:
                raise _lib.JogramopError('jg_scene_create: ' + L.jg_last_error().decode())
        prm = _lib.Params(threshold, query_distance, finger_open, chunk)
        self._e = L.jg_engine_create(self._r, self._s, self.device.index, C.byref(prm))
        if not self._e:
            raise _lib.JogramopError('jg_engine_create: ' + L.jg_last_error().decode())
        self.threshold, self.query_distance, self.finger_open = threshold, query_distance, finger_open
        if lanes != 1:
            self.set_lanes(lanes)

    def set_lanes(self, lanes):
        _lib.check(self._L.jg_engine_set_lanes(self._e, int(lanes)), 'jg_engine_set_lanes')

    @property
    def handle(self):
        """The ``jg_engine*`` as an integer: what ``torch.ops.jogramop.*`` (torch_ops.py) and C callers take."""
        return int(self._e)

    @property
    def pass_size(self):
        """Configurations one pass holds at most (larger calls run pass after pass on the engine's one arena)."""
        return int(self._L.jg_engine_pass_size(self._e))

    def close(self):
        L = self._L
        if getattr(self, '_e', None):
            L.jg_engine_free(self._e)
            self._e = None
        if getattr(self, '_s', None):
            L.jg_scene_free(self._s)
            self._s = None
        if getattr(self, '_r', None):
            L.jg_robot_free(self._r)
            self._r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _dev_f32(self, a, cols):
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=torch.float32)
        else:
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)
        t = t.reshape(-1, cols).contiguous()
        return t

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _outputs(self, n, want_dist, bits_out, dist_out):
        words = (n + 31) // 32
        if bits_out is not None:
            assert bits_out.dtype == torch.int32 and bits_out.is_contiguous() and bits_out.numel() >= words and \
                bits_out.device == self.device, 'bits_out must be a contiguous int32 device tensor of >= ceil(n/32) words'
            bits = bits_out[:words]
        else:
            bits = torch.empty((words,), device=self.device, dtype=torch.int32)
        if dist_out is not None:
            assert dist_out.dtype == torch.float32 and dist_out.is_contiguous() and dist_out.numel() >= n and \
                dist_out.device == self.device, 'dist_out must be a contiguous fp32 device tensor of >= n elements'
            dist, want_dist = dist_out[:n], True
        else:
            dist = torch.empty((n,), device=self.device, dtype=torch.float32) if want_dist else None
        return bits, dist

    # ------------------------------------------------------------------ queries
    def fk(self, q, link_ids=None, base_pose=None):
        """Link frames [n, len(link_ids), 3, 4] (fp32, device).  base_pose: 4x4 or 3x4 pre-multiplied pose."""
        q = self._dev_f32(q, self.dof)
        ids = _i32(range(self.robot.n_links) if link_ids is None else link_ids)
        out = torch.empty((q.shape[0], len(ids), 3, 4), device=self.device, dtype=torch.float32)
        bp = np.ascontiguousarray(_f64(base_pose)[:3, :]) if base_pose is not None else None
        with torch.cuda.device(self.device):
            _lib.check(self._L.jg_fk(self._e, q.data_ptr(), q.shape[0], self.dof, _p(ids), len(ids), _p(bp),
                                     out.data_ptr(), self._stream()), 'jg_fk')
        return out

    def check(self, q, flags=SCENE | PLANE, want_dist=True, want_reason=False, packed=False, bits_out=None, dist_out=None,
              reason_out=None):
        """-> (collides bool[n] (or packed int32 words), min_dist fp32[n] | None, reason uint8[n] | None).
        ``bits_out`` / ``dist_out``: preallocated device tensors the kernels write into directly (int32 words, at least
        ceil(n/32); fp32 [n]) -- e.g. this rank's slot of an all-gather buffer (``dist.GatherBuffer``), so that no copy
        sits between the finalize kernel and the collective."""
        q = self._dev_f32(q, self.dof)
        n = q.shape[0]
        bits, dist = self._outputs(n, want_dist, bits_out, dist_out)
        want_dist = dist is not None
        if reason_out is not None:
            assert reason_out.dtype == torch.uint8 and reason_out.is_contiguous() and reason_out.numel() >= n and \
                reason_out.device == self.device
            reason, want_reason = reason_out[:n], True
        else:
            reason = torch.empty((n,), device=self.device, dtype=torch.uint8) if want_reason else None
        with torch.cuda.device(self.device):
            _lib.check(self._L.jg_check(self._e, q.data_ptr(), n, self.dof, flags, bits.data_ptr(),
                                        dist.data_ptr() if want_dist else None,
                                        reason.data_ptr() if want_reason else None, self._stream()), 'jg_check')
        return (bits if packed else unpack_bits(bits, n)), dist, reason

    def capture_check(self, q, flags=SCENE | PLANE, want_dist=True, want_reason=False):
        """Capture one ``check`` of the device tensor ``q`` [n,dof] (n <= one pass) into a CUDA graph: the ~10 launches of
        a pass replay as ONE graph launch.  Returns (replay, bits, dist, reason): calling ``replay()`` re-runs the pass
        on the current contents of ``q`` and overwrites the returned output tensors (packed int32 words, fp32, uint8).
        Worth it for small batches issued over and over (a planner validating batch after batch), where launch gaps
        are a visible share of the pass."""
        assert isinstance(q, torch.Tensor) and q.is_cuda and q.dtype == torch.float32 and q.is_contiguous()
        n = q.shape[0]
        bits = torch.empty(((n + 31) // 32,), device=self.device, dtype=torch.int32)
        dist = torch.empty((n,), device=self.device, dtype=torch.float32) if want_dist else None
        reason = torch.empty((n,), device=self.device, dtype=torch.uint8) if want_reason else None

        def run():
            _lib.check(self._L.jg_check(self._e, q.data_ptr(), n, self.dof, flags, bits.data_ptr(),
                                        dist.data_ptr() if want_dist else None,
                                        reason.data_ptr() if want_reason else None, self._stream()), 'jg_check')
        with torch.cuda.device(self.device):
            run()                                   # grows the arena to its final size outside the capture
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run()
        return g.replay, bits, dist, reason

    def check_edges(self, qa, qb, substeps=64, flags=SCENE | PLANE, want_dist=True, packed=False, bits_out=None, dist_out=None):
        qa, qb = self._dev_f32(qa, self.dof), self._dev_f32(qb, self.dof)
        m = qa.shape[0]
        assert qb.shape[0] == m, 'qa and qb must hold the same number of edges'
        bits, dist = self._outputs(m, want_dist, bits_out, dist_out)
        want_dist = dist is not None
        with torch.cuda.device(self.device):
            _lib.check(self._L.jg_check_edges(self._e, qa.data_ptr(), qb.data_ptr(), m, self.dof, substeps, flags,
                                              bits.data_ptr(), dist.data_ptr() if want_dist else None,
                                              self._stream()), 'jg_check_edges')
        return (bits if packed else unpack_bits(bits, m)), dist

    def check_poses(self, poses, links, ee_link=None, flags=SCENE | PLANE, want_dist=True, packed=False, bits_out=None,
                    dist_out=None):
        """poses: [n,4,4] / [n,3,4] matrices or [n,7] (pos, quat xyzw) of link `ee_link` in the robot frame."""
        if isinstance(poses, torch.Tensor):
            shp = tuple(poses.shape)
        else:
            poses = np.asarray(poses)
            shp = poses.shape
        if len(shp) == 3 and shp[1:] == (4, 4):
            poses = poses[:, :3, :]
            fmt, cols = POSE_MAT34, 12
        elif len(shp) == 3 and shp[1:] == (3, 4):
            fmt, cols = POSE_MAT34, 12
        elif len(shp) == 2 and shp[1] == 7:
            fmt, cols = POSE_POS_QUAT, 7
        else:
            raise ValueError('poses must be [n,4,4], [n,3,4] or [n,7]')
        p = self._dev_f32(poses, cols)
        n = p.shape[0]
        links = _i32(links)
        ee = self.robot.end_effector_id if ee_link is None else ee_link
        bits, dist = self._outputs(n, want_dist, bits_out, dist_out)
        want_dist = dist is not None
        with torch.cuda.device(self.device):
            _lib.check(self._L.jg_check_poses(self._e, p.data_ptr(), n, fmt, ee, _p(links), len(links), flags,
                                              bits.data_ptr(), dist.data_ptr() if want_dist else None,
                                              self._stream()), 'jg_check_poses')
        return (bits if packed else unpack_bits(bits, n)), dist

    def check_host(self, q, flags=SCENE | PLANE, want_dist=True, want_reason=False, out=None):
        """End-to-end path with HOST buffers (numpy or CPU torch tensors, ideally pinned): H2D, kernels and D2H are
        pipelined inside the library.  Returns (bits uint32 words, dist, reason) as numpy views of `out` buffers."""
        if isinstance(q, torch.Tensor):
            assert q.device.type == 'cpu' and q.dtype == torch.float32 and q.is_contiguous()
            qn, qptr = q.shape[0], q.data_ptr()
        else:
            q = np.ascontiguousarray(q, dtype=np.float32)
            qn, qptr = q.shape[0], q.ctypes.data
        if out is None:
            out = {'bits': torch.empty(((qn + 31) // 32,), dtype=torch.int32, pin_memory=True),
                   'dist': torch.empty((qn,), dtype=torch.float32, pin_memory=True) if want_dist else None,
                   'reason': torch.empty((qn,), dtype=torch.uint8, pin_memory=True) if want_reason else None}
        _lib.check(self._L.jg_check_host(self._e, qptr, qn, self.dof, flags, out['bits'].data_ptr(),
                                         out['dist'].data_ptr() if out.get('dist') is not None else None,
                                         out['reason'].data_ptr() if out.get('reason') is not None else None),
                   'jg_check_host')
        return out['bits'], out.get('dist'), out.get('reason')

    def check_host_submit(self, q, flags=SCENE | PLANE, want_dist=True, want_reason=False, out=None):
        """Asynchronous ``check_host`` for streams of batches: returns a ticket at once; at most two batches may be in
        flight, so the copies of one batch overlap the kernels of the other.  ``q`` (and ``out``) must stay alive and
        unchanged until ``check_host_wait(ticket)`` returned; pinned memory avoids the staging copy."""
        if isinstance(q, torch.Tensor):
            assert q.device.type == 'cpu' and q.dtype == torch.float32 and q.is_contiguous()
            qn, qptr = q.shape[0], q.data_ptr()
        else:
            q = np.ascontiguousarray(q, dtype=np.float32)
            qn, qptr = q.shape[0], q.ctypes.data
        if out is None:
            out = {'bits': torch.empty(((qn + 31) // 32,), dtype=torch.int32, pin_memory=True),
                   'dist': torch.empty((qn,), dtype=torch.float32, pin_memory=True) if want_dist else None,
                   'reason': torch.empty((qn,), dtype=torch.uint8, pin_memory=True) if want_reason else None}
        t = C.c_int64(-1)
        _lib.check(self._L.jg_check_host_submit(self._e, qptr, qn, self.dof, flags, out['bits'].data_ptr(),
                                                out['dist'].data_ptr() if out.get('dist') is not None else None,
                                                out['reason'].data_ptr() if out.get('reason') is not None else None,
                                                C.byref(t)), 'jg_check_host_submit')
        return {'ticket': t.value, 'q': q, 'out': out}

    def check_host_wait(self, ticket):
        """Block until the batch of ``ticket`` is in its host buffers -> (bits uint32 words, dist, reason)."""
        _lib.check(self._L.jg_check_host_wait(self._e, ticket['ticket']), 'jg_check_host_wait')
        out = ticket['out']
        return out['bits'], out.get('dist'), out.get('reason')

    def ik(self, targets, q_rest, q_init=None, ee_link=None, iters=100, tol=1e-3, damping=0.05, null_gain=0.05,
           position_only=False):
        """Batched damped-least-squares IK.  targets: [n,4,4] / [n,3,4] poses of `ee_link` in the robot frame.
        -> (q [n,dof], err [n] = position error + deg/1000) as device tensors."""
        if isinstance(targets, torch.Tensor):
            t = targets
        else:
            t = torch.from_numpy(np.ascontiguousarray(targets, dtype=np.float32))
        t = t.to(device=self.device, dtype=torch.float32).reshape(-1, t.shape[-2], 4)[:, :3, :].contiguous()
        n = t.shape[0]
        qi = self._dev_f32(q_init, self.dof) if q_init is not None else None
        if qi is not None:
            assert qi.shape[0] == n, 'q_init must hold one start configuration per target'
        rest = _f64(q_rest)
        assert len(rest) == self.dof
        q = torch.empty((n, self.dof), device=self.device, dtype=torch.float32)
        err = torch.empty((n,), device=self.device, dtype=torch.float32)
        ee = self.robot.end_effector_id if ee_link is None else ee_link
        with torch.cuda.device(self.device):
            _lib.check(self._L.jg_ik(self._e, t.data_ptr(), n, ee, qi.data_ptr() if qi is not None else None, _p(rest),
                                     int(iters), float(tol), float(damping), float(null_gain), 1 if position_only else 0,
                                     q.data_ptr(), err.data_ptr(), self._stream()), 'jg_ik')
        return q, err

    def stats(self):
        st = _lib.Stats()
        _lib.check(self._L.jg_get_stats(self._e, C.byref(st)), 'jg_get_stats')
        return {f: getattr(st, f) for f, _ in st._fields_}

    def profile(self, enable=True):
        """Enable / disable per-kernel CUDA-event timing inside the library (resets the accumulators when enabling)."""
        _lib.check(self._L.jg_profile_enable(self._e, 1 if enable else 0), 'jg_profile_enable')

    def get_profile(self):
        """-> {'broadphase': (ms, launches), 'narrowphase': ..., 'penetration': ..., 'finalize': ...} since profile(True)."""
        ms = (C.c_double * 4)()
        n = (C.c_int64 * 4)()
        _lib.check(self._L.jg_get_profile(self._e, ms, n), 'jg_get_profile')
        return {k: (ms[i], n[i]) for i, k in enumerate(('broadphase', 'narrowphase', 'penetration', 'finalize'))}


class PoolTicket:
    """One batch in flight in an ``EnginePool``: ``event`` fires on the lane's stream when its kernels are done."""
    __slots__ = ('lane', 'event', 'result', 'keep')

    def __init__(self, lane, event, result, keep):
        self.lane, self.event, self.result, self.keep = lane, event, result, keep


class EnginePool:
    """``lanes`` engines (one scratch arena each) on ``lanes`` CUDA streams, handed whole batches round-robin.

    One pass is a chain of dependent kernels whose persistent narrow-phase stages end in tails (a polytope is a serial
    chain of ~10 expansions; the last ones of a stage run on mostly idle SMs) and which run at 13 warps per SM by
    shared-memory budget.  Nothing orders the passes of DIFFERENT batches, so with a second arena the next batch's
    kernels fill those tails and idle issue slots: measured 5.65e8 -> 6.3e8 (2 lanes) -> 6.7e8 checks/s (4 lanes) on
    scenario 011.  A lane's arena is sized by its largest batch (~16 KB per configuration for the Panda +
    scenario 011 tables), so four lanes of 1 M configurations hold ~62 GB of the 180 GB.

    ``submit`` returns at once; ``wait`` makes the CURRENT stream wait for that batch (no host synchronisation) and hands
    out its result tensors.  A C caller gets the same by creating one ``jg_engine`` per stream."""

    def __init__(self, robot, scene=None, device=0, lanes=4, **engine_kw):
        assert lanes >= 1
        self.engines = [CollisionEngine(robot, scene, device=device, **engine_kw) for _ in range(lanes)]
        self.device = self.engines[0].device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(lanes)]
        self.lanes = lanes
        self._next = 0
        self._host_inflight = [0] * lanes

    def close(self):
        for e in self.engines:
            e.close()
        self.engines = []

    @property
    def next_lane(self):
        """Index of the lane the next ``submit`` / ``submit_host`` will use."""
        return self._next

    def _lane(self):
        i = self._next
        self._next = (i + 1) % self.lanes
        return i

    def _submit_call(self, method, args, kw):
        i = self._lane()
        cur, st = torch.cuda.current_stream(self.device), self.streams[i]
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            res = getattr(self.engines[i], method)(*args, **kw)
            ev = torch.cuda.Event()
            ev.record(st)
        for t in args:
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(st)
        return PoolTicket(i, ev, res, args)

    def submit(self, q, **check_kw):
        """Queue ``CollisionEngine.check(q, **check_kw)`` on the next lane, ordered after everything the current stream
        holds so far (so a ``q`` produced there is ready) and after that lane's previous batch -> PoolTicket."""
        return self._submit_call('check', (q,), check_kw)

    def wait(self, ticket):
        """The current stream waits for the batch -> (collides, min_dist, reason) as ``check`` returns them."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ticket.event)
        for t in ticket.result:
            if t is not None:
                t.record_stream(cur)
        return ticket.result

    def drain(self):
        """The current stream waits for every lane."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)

    def check_many(self, batches, **check_kw):
        """All ``batches`` (a sequence of [n_i, dof] arrays / tensors) with up to ``lanes`` of them in flight -> list of
        ``check`` results in order."""
        tickets = [self.submit(q, **check_kw) for q in batches]
        return [self.wait(t) for t in tickets]

    # ---- ONE large call cut into pass-sized pieces that run side by side on the lanes, results in one set of tensors
    def _pieces(self, n, unit):
        """(offset, size) of the fewest equal pieces of at most ``unit`` (a multiple of 32: packed words stay aligned);
        a call that fits one pass is NOT cut (every pass pays ~0.5 ms of fixed cost: DESIGN.md 8)."""
        unit = max(32, unit // 32 * 32)
        k = max(1, -(-n // unit))
        piece = max(32, min(unit, (-(-n // k) + 31) // 32 * 32))
        return [(off, min(piece, n - off)) for off in range(0, n, piece)]

    def check(self, q, flags=SCENE | PLANE, want_dist=True, want_reason=False, packed=False, bits_out=None, dist_out=None):
        """``CollisionEngine.check`` for any n: the passes of a call larger than one pass run on different lanes (so an
        ``EnginePool`` can stand wherever an engine's ``check`` is called, e.g. ``dist.check_sharded``)."""
        e0 = self.engines[0]
        q = e0._dev_f32(q, e0.dof)
        n = q.shape[0]
        bits, dist = e0._outputs(n, want_dist, bits_out, dist_out)
        want_dist = dist is not None
        reason = torch.empty((n,), device=self.device, dtype=torch.uint8) if want_reason else None
        tks = [self._submit_call('check', (q[off:off + m],), dict(
            flags=flags, want_dist=want_dist, packed=True, bits_out=bits[off // 32:],
            dist_out=dist[off:] if want_dist else None, reason_out=reason[off:] if want_reason else None))
            for off, m in self._pieces(n, e0.pass_size)]
        for t in tks:
            self.wait(t)
        return (bits if packed else unpack_bits(bits, n)), dist, reason

    def check_edges(self, qa, qb, substeps=64, flags=SCENE | PLANE, want_dist=True, packed=False):
        """``CollisionEngine.check_edges`` with the passes (pass_size / substeps edges each) side by side on the lanes."""
        e0 = self.engines[0]
        qa, qb = e0._dev_f32(qa, e0.dof), e0._dev_f32(qb, e0.dof)
        m = qa.shape[0]
        assert qb.shape[0] == m, 'qa and qb must hold the same number of edges'
        bits = torch.empty(((m + 31) // 32,), device=self.device, dtype=torch.int32)
        dist = torch.empty((m,), device=self.device, dtype=torch.float32) if want_dist else None
        tks = [self._submit_call('check_edges', (qa[off:off + k], qb[off:off + k]), dict(
            substeps=substeps, flags=flags, want_dist=want_dist, packed=True, bits_out=bits[off // 32:],
            dist_out=dist[off:] if want_dist else None)) for off, k in self._pieces(m, e0.pass_size // max(substeps, 1))]
        for t in tks:
            self.wait(t)
        return (bits if packed else unpack_bits(bits, m)), dist

    def check_poses(self, poses, links, ee_link=None, flags=SCENE | PLANE, want_dist=True, packed=False):
        """``CollisionEngine.check_poses`` with the passes side by side on the lanes."""
        n = int(poses.shape[0])
        bits = torch.empty(((n + 31) // 32,), device=self.device, dtype=torch.int32)
        dist = torch.empty((n,), device=self.device, dtype=torch.float32) if want_dist else None
        tks = [self._submit_call('check_poses', (poses[off:off + k], links), dict(
            ee_link=ee_link, flags=flags, want_dist=want_dist, packed=True, bits_out=bits[off // 32:],
            dist_out=dist[off:] if want_dist else None)) for off, k in self._pieces(n, self.engines[0].pass_size)]
        for t in tks:
            self.wait(t)
        return (bits if packed else unpack_bits(bits, n)), dist

    # host-buffer form: lane i's engine pipelines upload / kernels / download itself (two tickets per engine)
    def submit_host(self, q, **kw):
        if self._host_inflight[self._next] >= 2:
            raise _lib.JogramopError('EnginePool.submit_host: more than 2 * lanes batches in flight; wait for the oldest first')
        i = self._lane()
        self._host_inflight[i] += 1
        return {'lane': i, 'inner': self.engines[i].check_host_submit(q, **kw), 'collected': False}

    def wait_host(self, ticket):
        res = self.engines[ticket['lane']].check_host_wait(ticket['inner'])
        if not ticket['collected']:
            ticket['collected'] = True
            self._host_inflight[ticket['lane']] -= 1
        return res


def measure_fp32_peak(device=0, reps=5):
    """FP32 FMA micro-benchmark on `device` -> TFLOP/s (roofline denominator of the narrow phase)."""
    out = C.c_double(0)
    _lib.check(_lib.load().jg_measure_fp32_peak(device, reps, C.byref(out)), 'jg_measure_fp32_peak')
    return out.value
