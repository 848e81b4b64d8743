"""Ingestion of the jogramop data formats into flat arrays for the collision engine.

Replaces what pyBullet's ``loadURDF`` / burg_toolkit's ``Scene.from_yaml`` +
``SimulatorBase.add_scene`` do for the reference (reference call sites:
``simulation.py:137-144`` robot URDF, ``scenario.py:24,64`` scene yaml, and the
obstacle merge of ``create_cpp_export.py:23-61``).  No pybullet / burg / open3d.

Semantics restated:
  * OBJ files are split into shapes on ``o`` / ``g`` statements; a file without
    groups is ONE shape; each shape becomes one convex hull (Bullet
    ``btConvexHullShape`` of the shape's vertices, collision margin 0.001).
  * ``<box>`` becomes a box whose 0.001 margin is *embedded* (core = box shrunk
    by the margin, ``btBoxShape``).
  * link index i = child link of joint i in URDF file order (depth first for a
    chain-like file); base = -1.
  * objects are placed with their mesh frame at ``pose`` and everything is
    expressed in the ROBOT frame: ``inv(robot_pose) @ pose @ v``
    (``create_cpp_export.py:36,60-61``); the ground plane z_world = 0 becomes
    ``n.p + d = 0`` with ``n = robot_pose[2,:3]``, ``d = robot_pose[2,3]``.

Everything here is host-side set-up code (one-time per scenario), numpy only.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

HULL_MARGIN = 0.001          # gUrdfDefaultCollisionMargin (Appendix C.1)
JOINT_FIXED, JOINT_REVOLUTE, JOINT_PRISMATIC = 0, 1, 2
SHAPE_HULL, SHAPE_BOX = 0, 1


# --------------------------------------------------------------------------- OBJ
def read_obj_shapes(path):
    """Parse a Wavefront OBJ into a list of shapes.

    Returns ``[(vertices[n,3] float64, triangles[m,3] int64 (indices local to the
    shape's vertex list))]``.  Handles ``v``, ``f a b c``, ``f a/b/c``, ``f a//c``,
    polygons (fan triangulation), negative indices; ignores vt/vn/usemtl/s/l/mtllib.
    Shapes are split on ``o`` and ``g`` (one convex piece per group); a file with no
    groups is one shape (pyBullet reading, Appendix C.1 / quirk Q1).
    """
    all_v = []
    shapes = []          # list of face lists (global vertex ids)
    cur = None
    with open(path, 'r', errors='replace') as fh:
        for line in fh:
            if not line or line[0] == '#':
                continue
            tok = line.split()
            if not tok:
                continue
            key = tok[0]
            if key == 'v':
                all_v.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif key in ('o', 'g'):
                cur = []
                shapes.append(cur)
            elif key == 'f':
                if cur is None:
                    cur = []
                    shapes.append(cur)
                ids = []
                for t in tok[1:]:
                    i = int(t.split('/')[0])
                    ids.append(i - 1 if i > 0 else len(all_v) + i)
                for k in range(1, len(ids) - 1):
                    cur.append((ids[0], ids[k], ids[k + 1]))
    all_v = np.asarray(all_v, dtype=np.float64).reshape(-1, 3)
    out = []
    for faces in shapes:
        if not faces:
            continue
        f = np.asarray(faces, dtype=np.int64)
        used, inv = np.unique(f.reshape(-1), return_inverse=True)
        out.append((all_v[used], inv.reshape(-1, 3)))
    if not out and len(all_v):
        out.append((all_v, np.zeros((0, 3), dtype=np.int64)))
    return out


def convex_hull_vertices(points):
    """Vertices of the convex hull of ``points`` (float64 [n,3]) -> [h,3].

    Support functions (all the engine needs) are unchanged by dropping interior or
    duplicated points, so this is equivalent to Bullet's hull of *all* shape vertices.
    """
    from scipy.spatial import ConvexHull, QhullError
    pts = np.unique(np.asarray(points, dtype=np.float64), axis=0)
    if len(pts) <= 4:
        return pts
    try:
        hull = ConvexHull(pts)
        return pts[np.sort(hull.vertices)]
    except QhullError:
        return pts  # degenerate (flat) shape: keep all points, support is still exact


def split_connected_components(vertices, triangles, tol=1e-6):
    """Split a triangle mesh into connected components after welding vertices that
    coincide within ``tol`` (used for the stale ``gate_0_54`` decomposition, quirk Q1)."""
    key = np.round(vertices / tol).astype(np.int64)
    _, weld = np.unique(key, axis=0, return_inverse=True)
    weld = weld.reshape(-1)
    parent = np.arange(weld.max() + 1)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for tri in triangles:
        a, b, c = (find(weld[i]) for i in tri)
        parent[b] = a
        parent[find(c)] = a
    roots = np.array([find(weld[t[0]]) for t in triangles])
    comps = []
    _, first = np.unique(roots, return_index=True)
    for r in roots[np.sort(first)]:      # components in order of first appearance
        tri = triangles[roots == r]
        used, inv = np.unique(tri.reshape(-1), return_inverse=True)
        comps.append((vertices[used], inv.reshape(-1, 3)))
    return comps


# --------------------------------------------------------------------------- math
def rpy_to_matrix(rpy):
    """URDF fixed-axis XYZ: R = Rz(yaw) Ry(pitch) Rx(roll) (Appendix A)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr]])


def make_tf(R=None, t=None):
    T = np.eye(4)
    if R is not None:
        T[:3, :3] = R
    if t is not None:
        T[:3, 3] = t
    return T


def _origin(elem):
    T = np.eye(4)
    if elem is None:
        return T
    o = elem.find('origin')
    if o is None:
        return T
    xyz = [float(x) for x in o.get('xyz', '0 0 0').split()]
    rpy = [float(x) for x in o.get('rpy', '0 0 0').split()]
    return make_tf(rpy_to_matrix(rpy), xyz)


# --------------------------------------------------------------------------- robot
@dataclass
class RobotModel:
    """Flat description of a URDF robot: kinematic chain + convex collision shapes.

    Link i is the child of joint i (pyBullet numbering).  Shape vertices are stored
    in the LINK frame (the URDF ``<collision><origin>`` is pre-applied).
    """
    name: str
    link_names: list
    joint_names: list
    joint_type: np.ndarray        # [L] int32
    joint_parent: np.ndarray      # [L] int32, parent link index, -1 = base
    joint_origin: np.ndarray      # [L,4,4] float64 (parent link frame -> joint frame)
    joint_axis: np.ndarray        # [L,3] float64
    joint_limits: np.ndarray      # [L,2] float64 (0,0 for fixed)
    shape_link: np.ndarray        # [S] int32
    shape_type: np.ndarray        # [S] int32  SHAPE_HULL / SHAPE_BOX
    shape_margin: np.ndarray      # [S] float64 collision margin subtracted from the core distance
    shape_vert_offset: np.ndarray  # [S+1] int32 into core_verts
    core_verts: np.ndarray        # [V,3] float64 GJK core vertices (hull verts; box shrunk by margin)
    shape_plane_offset: np.ndarray  # [S+1] into plane_verts
    plane_verts: np.ndarray       # [W,3] vertices used for the plane test (hull: core; box: exact corners)
    shape_plane_margin: np.ndarray  # [S] margin subtracted in the plane test (hull 0.001, box 0)
    arm_joint_ids: list = field(default_factory=list)
    finger_joint_ids: list = field(default_factory=list)
    end_effector_id: int = -1
    self_pairs: np.ndarray = None  # [K,2] int32 link index pairs (simulation.py:439-464)

    @property
    def n_links(self):
        return len(self.link_names)

    @property
    def n_shapes(self):
        return len(self.shape_link)

    def shape_verts(self, s):
        return self.core_verts[self.shape_vert_offset[s]:self.shape_vert_offset[s + 1]]

    def shape_plane_verts(self, s):
        return self.plane_verts[self.shape_plane_offset[s]:self.shape_plane_offset[s + 1]]

    def to_dict(self):
        d = {k: getattr(self, k) for k in (
            'joint_type', 'joint_parent', 'joint_origin', 'joint_axis', 'joint_limits', 'shape_link',
            'shape_type', 'shape_margin', 'shape_vert_offset', 'core_verts', 'shape_plane_offset',
            'plane_verts', 'shape_plane_margin', 'self_pairs')}
        d['name'] = np.array(self.name)
        d['link_names'] = np.array(self.link_names)
        d['joint_names'] = np.array(self.joint_names)
        d['arm_joint_ids'] = np.asarray(self.arm_joint_ids, dtype=np.int32)
        d['finger_joint_ids'] = np.asarray(self.finger_joint_ids, dtype=np.int32)
        d['end_effector_id'] = np.array(self.end_effector_id, dtype=np.int32)
        return d

    @staticmethod
    def from_dict(d):
        return RobotModel(
            name=str(d['name']), link_names=[str(x) for x in d['link_names']],
            joint_names=[str(x) for x in d['joint_names']],
            joint_type=d['joint_type'], joint_parent=d['joint_parent'], joint_origin=d['joint_origin'],
            joint_axis=d['joint_axis'], joint_limits=d['joint_limits'], shape_link=d['shape_link'],
            shape_type=d['shape_type'], shape_margin=d['shape_margin'],
            shape_vert_offset=d['shape_vert_offset'], core_verts=d['core_verts'],
            shape_plane_offset=d['shape_plane_offset'], plane_verts=d['plane_verts'],
            shape_plane_margin=d['shape_plane_margin'],
            arm_joint_ids=[int(x) for x in d['arm_joint_ids']],
            finger_joint_ids=[int(x) for x in d['finger_joint_ids']],
            end_effector_id=int(d['end_effector_id']), self_pairs=d['self_pairs'])

    def save(self, fn):
        np.savez_compressed(fn, **self.to_dict())

    @staticmethod
    def load(fn):
        with np.load(fn, allow_pickle=False) as z:
            return RobotModel.from_dict({k: z[k] for k in z.files})


def self_collision_pairs(with_platform):
    """Ordered link pairs of ``FrankaRobot.in_self_collision`` (simulation.py:439-464)."""
    if with_platform:
        max_link, ignore, first_links = 14, [0, 1, 10, 14], [2, 3, 4, 5, 6, 7, 8]
    else:
        max_link, ignore, first_links = 11, [7, 11], [1, 2, 3, 4, 5]
    pairs = []
    for f in first_links:
        for c in range(f + 2, max_link + 1):
            if c not in ignore:
                pairs.append((f, c))
    return np.asarray(pairs, dtype=np.int32)


def load_robot_urdf(urdf_path, with_platform=None, margin=HULL_MARGIN):
    """Read a jogramop Panda URDF (``mobile_panda.urdf`` / ``panda.urdf``) into a RobotModel."""
    root = ET.parse(urdf_path).getroot()
    base_dir = os.path.dirname(os.path.abspath(urdf_path))
    links = {l.get('name'): l for l in root.findall('link')}
    joints = root.findall('joint')
    children = {j.find('child').get('link') for j in joints}
    base = [n for n in links if n not in children]
    assert len(base) == 1, 'expected a single root link'
    # pyBullet numbering: depth-first from the base, children in file order.
    by_parent = {}
    for j in joints:
        by_parent.setdefault(j.find('parent').get('link'), []).append(j)
    order = []

    def visit(link):
        for j in by_parent.get(link, []):
            order.append(j)
            visit(j.find('child').get('link'))
    visit(base[0])

    link_index = {base[0]: -1}
    for i, j in enumerate(order):
        link_index[j.find('child').get('link')] = i
    L = len(order)
    jt = np.zeros(L, np.int32)
    jp = np.zeros(L, np.int32)
    jo = np.zeros((L, 4, 4))
    ja = np.zeros((L, 3))
    jl = np.zeros((L, 2))
    for i, j in enumerate(order):
        t = j.get('type')
        jt[i] = {'fixed': JOINT_FIXED, 'revolute': JOINT_REVOLUTE, 'continuous': JOINT_REVOLUTE,
                 'prismatic': JOINT_PRISMATIC}[t]
        jp[i] = link_index[j.find('parent').get('link')]
        jo[i] = _origin(j)
        ax = j.find('axis')
        if ax is not None and jt[i] != JOINT_FIXED:
            ja[i] = [float(x) for x in ax.get('xyz').split()]
        elif jt[i] != JOINT_FIXED:
            ja[i] = [1.0, 0.0, 0.0]
        lim = j.find('limit')
        if lim is not None and jt[i] != JOINT_FIXED:
            jl[i] = [float(lim.get('lower', 0)), float(lim.get('upper', 0))]

    s_link, s_type, s_margin, s_pm = [], [], [], []
    core, plane = [], []
    # link -1 is the base link (panda.urdf: panda_link0 carries a collision mesh)
    for i in range(-1, L):
        link = links[base[0] if i < 0 else order[i].find('child').get('link')]
        for col in link.findall('collision'):
            T = _origin(col)
            geom = col.find('geometry')
            mesh, box = geom.find('mesh'), geom.find('box')
            if mesh is not None:
                scale = [float(x) for x in mesh.get('scale', '1 1 1').split()]
                for v, _ in read_obj_shapes(os.path.join(base_dir, mesh.get('filename'))):
                    h = convex_hull_vertices(v * scale)
                    h = h @ T[:3, :3].T + T[:3, 3]
                    s_link.append(i); s_type.append(SHAPE_HULL); s_margin.append(margin); s_pm.append(margin)
                    core.append(h); plane.append(h)
            elif box is not None:
                half = 0.5 * np.array([float(x) for x in box.get('size').split()])
                sg = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)
                c = (sg * (half - margin)) @ T[:3, :3].T + T[:3, 3]
                e = (sg * half) @ T[:3, :3].T + T[:3, 3]
                s_link.append(i); s_type.append(SHAPE_BOX); s_margin.append(margin); s_pm.append(0.0)
                core.append(c); plane.append(e)
            else:
                raise NotImplementedError('only <mesh> and <box> collision geometry is supported')
    off = np.cumsum([0] + [len(c) for c in core]).astype(np.int32)
    poff = np.cumsum([0] + [len(c) for c in plane]).astype(np.int32)

    names = [j.find('child').get('link') for j in order]
    if with_platform is None:
        with_platform = 'mobile_platform' in names
    if with_platform:   # simulation.py:156-160
        arm, fingers, ee = [0, 1, 3, 4, 5, 6, 7, 8, 9], [12, 13], 14
    else:               # simulation.py:161-164
        arm, fingers, ee = [0, 1, 2, 3, 4, 5, 6], [9, 10], 11
    return RobotModel(
        name=os.path.basename(urdf_path), link_names=names, joint_names=[j.get('name') for j in order],
        joint_type=jt, joint_parent=jp, joint_origin=jo, joint_axis=ja, joint_limits=jl,
        shape_link=np.asarray(s_link, np.int32), shape_type=np.asarray(s_type, np.int32),
        shape_margin=np.asarray(s_margin), shape_vert_offset=off, core_verts=np.concatenate(core),
        shape_plane_offset=poff, plane_verts=np.concatenate(plane), shape_plane_margin=np.asarray(s_pm),
        arm_joint_ids=arm, finger_joint_ids=fingers, end_effector_id=ee,
        self_pairs=self_collision_pairs(with_platform))


# --------------------------------------------------------------------------- scene
@dataclass
class SceneObject:
    """Stand-in for burg_toolkit's ObjectInstance: ``object_type`` name + 4x4 ``pose`` (world)."""
    object_type: str
    pose: np.ndarray
    vhacd_fn: str = None
    is_target: bool = False
    mesh_fn: str = None          # the library's ``mesh_fn``: what burg's ``ObjectInstance.get_mesh()`` returns


@dataclass
class SceneDescription:
    """Stand-in for burg_toolkit's ``Scene``: ``objects`` (targets) and ``bg_objects`` (obstacles)."""
    objects: list
    bg_objects: list
    ground_area: tuple = (2.0, 2.0)
    scene_fn: str = None


@dataclass
class SceneModel:
    """Obstacle geometry of one scenario in the ROBOT frame.

    ``pieces``: convex pieces (hull vertices) — the pyBullet collision model;
    ``tri_verts``/``tris``: the merged triangle mesh — identical to the reference's
    ``export/obstacles.obj`` (create_cpp_export.py:40-61), objects first, then bg_objects.
    """
    name: str
    robot_pose: np.ndarray           # [4,4]
    piece_vert_offset: np.ndarray    # [P+1] int32
    piece_verts: np.ndarray          # [V,3] float64
    piece_body: np.ndarray           # [P] int32: index into body_names
    piece_margin: np.ndarray         # [P] float64
    body_names: list                 # plane is NOT in this list
    body_is_target: np.ndarray       # [B] bool
    tri_verts: np.ndarray            # [T,3] float64
    tris: np.ndarray                 # [F,3] int32
    tri_body: np.ndarray             # [F] int32
    plane: np.ndarray                # [4] (n, d): signed distance of robot-frame point p = n.p + d (world z=0)
    has_plane: bool = True

    @property
    def n_pieces(self):
        return len(self.piece_body)

    def piece(self, p):
        return self.piece_verts[self.piece_vert_offset[p]:self.piece_vert_offset[p + 1]]

    def retarget(self, robot_pose):
        """The same scene expressed in the frame of a robot placed at ``robot_pose`` (4x4, world)."""
        robot_pose = np.asarray(robot_pose, dtype=np.float64)
        if np.array_equal(robot_pose, self.robot_pose):
            return self
        T = np.linalg.inv(robot_pose) @ self.robot_pose
        import copy
        m = copy.copy(self)
        m.robot_pose = robot_pose
        m.piece_verts = self.piece_verts @ T[:3, :3].T + T[:3, 3]
        m.tri_verts = self.tri_verts @ T[:3, :3].T + T[:3, 3]
        m.plane = np.array([robot_pose[2, 0], robot_pose[2, 1], robot_pose[2, 2], robot_pose[2, 3]])
        return m

    def subset(self, bodies, with_plane):
        """Scene restricted to the pieces of the given body indices (and optionally the plane)."""
        import copy
        keep = [p for p in range(self.n_pieces) if self.piece_body[p] in set(bodies)]
        m = copy.copy(self)
        m.piece_verts = (np.concatenate([self.piece(p) for p in keep]) if keep else np.zeros((0, 3)))
        m.piece_vert_offset = np.cumsum([0] + [len(self.piece(p)) for p in keep]).astype(np.int32)
        m.piece_body = self.piece_body[keep]
        m.piece_margin = self.piece_margin[keep]
        m.has_plane = bool(with_plane and self.has_plane)
        # the merged triangle mesh follows the same body selection (mesh-mode engines)
        sel = np.isin(self.tri_body, list(set(bodies)))
        tris = self.tris[sel]
        used, inv = np.unique(tris.reshape(-1), return_inverse=True)
        m.tri_verts = self.tri_verts[used] if len(used) else np.zeros((0, 3))
        m.tris = inv.reshape(-1, 3).astype(np.int32)
        m.tri_body = self.tri_body[sel]
        return m

    def to_dict(self):
        d = {k: getattr(self, k) for k in (
            'robot_pose', 'piece_vert_offset', 'piece_verts', 'piece_body', 'piece_margin',
            'body_is_target', 'tri_verts', 'tris', 'tri_body')}
        d['name'] = np.array(self.name)
        d['body_names'] = np.array(self.body_names)
        d['plane'] = np.asarray(self.plane, dtype=np.float64)
        d['has_plane'] = np.array(self.has_plane)
        return d

    @staticmethod
    def from_dict(d):
        return SceneModel(
            name=str(d['name']), robot_pose=d['robot_pose'], piece_vert_offset=d['piece_vert_offset'],
            piece_verts=d['piece_verts'], piece_body=d['piece_body'], piece_margin=d['piece_margin'],
            body_names=[str(x) for x in d['body_names']], body_is_target=d['body_is_target'],
            tri_verts=d['tri_verts'], tris=d['tris'], tri_body=d['tri_body'],
            plane=d['plane'], has_plane=bool(d['has_plane']))

    def save(self, fn):
        np.savez_compressed(fn, **self.to_dict())

    @staticmethod
    def load(fn):
        with np.load(fn, allow_pickle=False) as z:
            return SceneModel.from_dict({k: z[k] for k in z.files})


def load_object_library(lib_fn):
    import yaml
    with open(lib_fn) as fh:
        lib = yaml.safe_load(fh)
    base = os.path.dirname(os.path.abspath(lib_fn))
    out = {}
    for o in lib['objects']:
        out[o['identifier']] = {
            'vhacd_fn': os.path.join(base, o['vhacd_fn']) if o.get('vhacd_fn') else None,
            'urdf_fn': os.path.join(base, o['urdf_fn']) if o.get('urdf_fn') else None,
            'mesh_fn': os.path.join(base, o['mesh_fn']) if o.get('mesh_fn') else None,
            'scale': o.get('scale', 1.0),
        }
    return out


def load_scene_yaml(scene_fn):
    """``burg.Scene.from_yaml`` stand-in (scenario.py:24): returns (SceneDescription, library dict)."""
    import yaml
    with open(scene_fn) as fh:
        y = yaml.safe_load(fh)
    lib_fn = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(scene_fn)), y['object_library_fn']))
    lib = load_object_library(lib_fn)

    def inst(e, is_target):
        t = e['object_type']
        if t not in lib:
            raise KeyError(f'object type {t} not in object library {lib_fn}')
        return SceneObject(t, np.asarray(e['pose'], dtype=np.float64), lib[t]['vhacd_fn'], is_target, lib[t]['mesh_fn'])
    scene = SceneDescription(
        objects=[inst(e, True) for e in (y.get('objects') or [])],
        bg_objects=[inst(e, False) for e in (y.get('bg_objects') or [])],
        ground_area=(y.get('ground_area_x', 2), y.get('ground_area_y', 2)), scene_fn=scene_fn)
    for o in scene.objects + scene.bg_objects:
        if lib[o.object_type]['scale'] not in (None, 1.0):
            # create_cpp_export.py:33-34
            raise NotImplementedError('did not implement scale feature for VHACD export')
    return scene, lib


def build_scene_model(scene, robot_pose, name='scene', ungrouped_obj_policy='single_hull',
                      piece_overrides=None, margin=HULL_MARGIN, with_plane=True):
    """Scene description -> SceneModel in the robot frame.

    Body order follows ``create_cpp_export.py:41``: objects first, then bg_objects (this is
    also the triangle order of the exported ``obstacles.obj``).  ``ungrouped_obj_policy``:
    ``single_hull`` (pyBullet-faithful) or ``connected_components`` (quirk Q1).
    ``piece_overrides`` maps object_type -> list of (vertices, triangles) (object frame) to use instead
    of the file's shapes (used for the recovered ``gate_0_54`` decomposition).
    """
    robot_pose = np.asarray(robot_pose, dtype=np.float64)
    to_robot = np.linalg.inv(robot_pose)
    pv, pb, tv, tf_, tb, names, is_t = [], [], [], [], [], [], []
    n_tv = 0
    for b, obj in enumerate(list(scene.objects) + list(scene.bg_objects)):
        names.append(obj.object_type)
        is_t.append(bool(obj.is_target))
        T = to_robot @ obj.pose
        if piece_overrides and obj.object_type in piece_overrides:
            shapes = [(np.asarray(v, float), np.asarray(f, np.int64)) for v, f in piece_overrides[obj.object_type]]
        else:
            shapes = read_obj_shapes(obj.vhacd_fn)
        if len(shapes) == 1 and ungrouped_obj_policy == 'connected_components':
            hull_sets = [v for v, _ in split_connected_components(*shapes[0])]
        else:
            hull_sets = [v for v, _ in shapes]
        for v in hull_sets:
            h = convex_hull_vertices(v)
            pv.append(h @ T[:3, :3].T + T[:3, 3])
            pb.append(b)
        for v, f in shapes:      # merged triangle mesh (all vertices kept, like open3d's reader)
            tv.append(v @ T[:3, :3].T + T[:3, 3])
            tf_.append(f + n_tv)
            tb.append(np.full(len(f), b, np.int32))
            n_tv += len(v)
    off = np.cumsum([0] + [len(p) for p in pv]).astype(np.int32)
    return SceneModel(
        name=name, robot_pose=robot_pose, piece_vert_offset=off,
        piece_verts=np.concatenate(pv) if pv else np.zeros((0, 3)),
        piece_body=np.asarray(pb, np.int32), piece_margin=np.full(len(pb), margin),
        body_names=names, body_is_target=np.asarray(is_t, bool),
        tri_verts=np.concatenate(tv) if tv else np.zeros((0, 3)),
        tris=(np.concatenate(tf_) if tf_ else np.zeros((0, 3))).astype(np.int32),
        tri_body=np.concatenate(tb) if tb else np.zeros(0, np.int32),
        plane=np.array([robot_pose[2, 0], robot_pose[2, 1], robot_pose[2, 2], robot_pose[2, 3]]),
        has_plane=with_plane)


def merge_meshes(bodies, robot_pose, name='visual'):
    """``bodies``: [(vertices [n,3] in the object's mesh frame, triangles [m,3], pose 4x4 world, is_target)] -> a SceneModel
    that carries ONLY the merged triangle mesh (robot frame) and the plane: the scene as burg's
    ``AntipodalGraspSampler.check_collisions`` sees it (``obj.get_mesh()`` of every object and bg object: the library's
    ``mesh_fn`` meshes, not the VHACD pieces; create_graspset.py:54-59).  For mesh-mode engines."""
    robot_pose = np.asarray(robot_pose, dtype=np.float64)
    to_robot = np.linalg.inv(robot_pose)
    tv, tf_, tb, is_t, names = [], [], [], [], []
    n_tv = 0
    for b, (v, f, pose, target) in enumerate(bodies):
        T = to_robot @ np.asarray(pose, dtype=np.float64)
        tv.append(np.asarray(v, dtype=np.float64) @ T[:3, :3].T + T[:3, 3])
        tf_.append(np.asarray(f, dtype=np.int64) + n_tv)
        tb.append(np.full(len(f), b, np.int32))
        n_tv += len(v)
        is_t.append(bool(target))
        names.append(f'body{b}')
    return SceneModel(
        name=name, robot_pose=robot_pose, piece_vert_offset=np.zeros(1, np.int32), piece_verts=np.zeros((0, 3)),
        piece_body=np.zeros(0, np.int32), piece_margin=np.zeros(0), body_names=names,
        body_is_target=np.asarray(is_t, bool), tri_verts=np.concatenate(tv) if tv else np.zeros((0, 3)),
        tris=(np.concatenate(tf_) if tf_ else np.zeros((0, 3))).astype(np.int32),
        tri_body=np.concatenate(tb) if tb else np.zeros(0, np.int32),
        plane=np.array([robot_pose[2, 0], robot_pose[2, 1], robot_pose[2, 2], robot_pose[2, 3]]), has_plane=True)


def read_obj_mesh(path):
    """Whole OBJ file as one welded-as-written triangle mesh (all groups merged): (vertices, triangles)."""
    shapes = read_obj_shapes(path)
    v = np.concatenate([s[0] for s in shapes])
    off = np.cumsum([0] + [len(s[0]) for s in shapes])
    f = np.concatenate([s[1] + o for s, o in zip(shapes, off)])
    return v, f


def default_robot_pose():
    """scenario.py:69-80."""
    return np.array([[0.0, -1.0, 0.0, 1.0], [1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.05], [0.0, 0.0, 0.0, 1.0]])


# --------------------------------------------------------------------------- writers (hand a model to file-based callers)
def _hull_faces(v):
    """Triangles of the convex hull of v [n,3] (indices into v); a flat / tiny set gets a fan so every vertex is used."""
    from scipy.spatial import ConvexHull, QhullError
    if len(v) >= 4:
        try:
            return ConvexHull(v).simplices
        except QhullError:
            pass
    return np.array([[0, i, i + 1] for i in range(1, max(len(v) - 1, 2))]) % max(len(v), 1)


def _matrix_to_rpy(R):
    """Inverse of rpy_to_matrix (URDF fixed-axis XYZ)."""
    pitch = -np.arcsin(np.clip(R[2, 0], -1.0, 1.0))
    if abs(np.cos(pitch)) > 1e-9:
        return np.arctan2(R[2, 1], R[2, 2]), pitch, np.arctan2(R[1, 0], R[0, 0])
    return np.arctan2(-R[1, 2], R[1, 1]), pitch, 0.0


def export_robot_urdf(model, out_dir, name='robot'):
    """Write a RobotModel as ``<out_dir>/<name>.urdf`` + one OBJ per hull shape (``meshes/shape_<s>.obj``: the hull
    vertices and the hull's triangles), so that file-based callers -- ``jg_robot_load`` of the C-ABI, the sister C++
    planners -- get exactly this collision model.  Boxes are written as ``<box>`` (they must be axis-aligned in their
    link frame, as the mobile platform is).  Returns the URDF path."""
    os.makedirs(os.path.join(out_dir, 'meshes'), exist_ok=True)
    base_name = 'base_link_export'
    lines = ['<?xml version="1.0" ?>', f'<robot name="{name}">']

    def collisions(link):
        out = []
        for s in np.nonzero(model.shape_link == link)[0]:
            if model.shape_type[s] == SHAPE_BOX:
                e = model.shape_plane_verts(s)
                lo, hi = e.min(0), e.max(0)
                if len(e) != 8 or not np.allclose(np.sort(np.unique(np.round(e, 12), axis=0), axis=0).shape, (8, 3)):
                    raise NotImplementedError('only axis-aligned boxes can be exported')
                c, size = [float(x) for x in 0.5 * (lo + hi)], [float(x) for x in hi - lo]
                out.append(f'    <collision><origin xyz="{c[0]!r} {c[1]!r} {c[2]!r}" rpy="0 0 0"/>'
                           f'<geometry><box size="{size[0]!r} {size[1]!r} {size[2]!r}"/></geometry></collision>')
            else:
                v = model.shape_verts(s)
                fn = os.path.join('meshes', f'shape_{s}.obj')
                with open(os.path.join(out_dir, fn), 'w') as fh:
                    fh.write(f'o shape_{s}\n')
                    for p in v:
                        fh.write(f'v {float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n')
                    for t in _hull_faces(v):
                        fh.write(f'f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n')
                out.append(f'    <collision><origin xyz="0 0 0" rpy="0 0 0"/><geometry><mesh filename="{fn}"/></geometry></collision>')
        return out

    lines += [f'  <link name="{base_name}">'] + collisions(-1) + ['  </link>']
    for i, ln in enumerate(model.link_names):
        lines += [f'  <link name="{ln}">'] + collisions(i) + ['  </link>']
    types = {JOINT_FIXED: 'fixed', JOINT_REVOLUTE: 'revolute', JOINT_PRISMATIC: 'prismatic'}
    for i, jn in enumerate(model.joint_names):
        T = model.joint_origin[i]
        r, p, y = _matrix_to_rpy(T[:3, :3])
        parent = base_name if model.joint_parent[i] < 0 else model.link_names[model.joint_parent[i]]
        lines.append(f'  <joint name="{jn}" type="{types[int(model.joint_type[i])]}">')
        lines.append(f'    <origin xyz="{float(T[0, 3])!r} {float(T[1, 3])!r} {float(T[2, 3])!r}" rpy="{float(r)!r} {float(p)!r} {float(y)!r}"/>')
        lines.append(f'    <parent link="{parent}"/><child link="{model.link_names[i]}"/>')
        if model.joint_type[i] != JOINT_FIXED:
            a = model.joint_axis[i]
            lines.append(f'    <axis xyz="{float(a[0])!r} {float(a[1])!r} {float(a[2])!r}"/>')
            lines.append(f'    <limit lower="{float(model.joint_limits[i, 0])!r}" upper="{float(model.joint_limits[i, 1])!r}" effort="0" velocity="0"/>')
        lines.append('  </joint>')
    lines.append('</robot>')
    path = os.path.join(out_dir, f'{name}.urdf')
    with open(path, 'w') as fh:
        fh.write('\n'.join(lines) + '\n')
    return path


def export_scene_files(scene_model, out_dir, scenario='000'):
    """Write a SceneModel as ``scenarios/<scenario>/scene.yaml`` + ``object_library/object_library.yaml`` + one OBJ per
    body (its convex pieces as ``o`` groups, robot-frame coordinates; every body is placed with pose = robot_pose, so
    loading with that robot pose reproduces the model).  Returns the scene.yaml path."""
    lib_dir = os.path.join(out_dir, 'object_library')
    sc_dir = os.path.join(out_dir, 'scenarios', str(scenario))
    os.makedirs(os.path.join(lib_dir, 'vhacd'), exist_ok=True)
    os.makedirs(sc_dir, exist_ok=True)
    lib = ['description: exported by jogramop_framework_b200.ingest.export_scene_files', 'name: export', 'objects:']
    entries = {True: [], False: []}
    for b, (bname, is_t) in enumerate(zip(scene_model.body_names, scene_model.body_is_target)):
        ident = f'body{b}_{bname}'
        fn = os.path.join('vhacd', f'{ident}.obj')
        n_v = 0
        with open(os.path.join(lib_dir, fn), 'w') as fh:
            for p in np.nonzero(scene_model.piece_body == b)[0]:
                v = scene_model.piece(p)
                fh.write(f'o piece_{p}\n')
                for x in v:
                    fh.write(f'v {float(x[0])!r} {float(x[1])!r} {float(x[2])!r}\n')
                for t in _hull_faces(v):
                    fh.write(f'f {t[0] + 1 + n_v} {t[1] + 1 + n_v} {t[2] + 1 + n_v}\n')
                n_v += len(v)
        lib += [f'- identifier: {ident}', f'  mesh_fn: {fn}', f'  name: {ident}', '  scale: 1.0', '  thumbnail_fn: null',
                f'  vhacd_fn: {fn}']
        pose = ['  pose:'] + [('  - - ' if c == 0 else '    - ') + repr(float(scene_model.robot_pose[r, c]))
                              for r in range(4) for c in range(4)]
        entries[bool(is_t)] += [f'- object_type: {ident}'] + pose
    with open(os.path.join(lib_dir, 'object_library.yaml'), 'w') as fh:
        fh.write('\n'.join(lib) + '\n')
    y = ['bg_objects:' + ('' if entries[False] else ' []')] + entries[False] + ['ground_area_x: 2', 'ground_area_y: 2',
         'object_library_fn: ../../object_library/object_library.yaml', 'objects:' + ('' if entries[True] else ' []')] + \
        entries[True] + ['printout: null', 'yaml_file_type: Scene', "yaml_file_version: '1.0'"]
    path = os.path.join(sc_dir, 'scene.yaml')
    with open(path, 'w') as fh:
        fh.write('\n'.join(y) + '\n')
    return path
