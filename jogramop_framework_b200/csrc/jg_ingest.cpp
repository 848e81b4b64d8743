// jg_ingest.cpp -- file-level loaders of the C-ABI: jg_robot_load / jg_scene_load / jg_scene_load_mesh.
//
// What pybullet.loadURDF + burg's Scene.from_yaml / SimulatorBase.add_scene do for the reference (simulation.py:137-144,
// scenario.py:24,64, obstacle merge of create_cpp_export.py:23-61), for callers that have no Python: the sister C++
// planners (README.md:52-56) can load robots/franka_panda/mobile_panda.urdf and scenarios/<id>/scene.yaml directly.
// Same semantics as jogramop_framework_b200/ingest.py:
//   * OBJ files split into shapes on `o` / `g`; a file without groups is ONE shape; every shape = convex hull of its
//     vertices with a 0.001 collision margin; <box> embeds its margin (core = box shrunk by the margin);
//   * link i = child link of joint i, depth first from the root, children in file order; base = -1;
//   * scene objects are placed with their mesh frame at `pose` and expressed in the ROBOT frame,
//     inv(robot_pose) @ pose @ v, objects first, then bg_objects; ground plane z_world = 0.
// Self-contained readers (no third-party code): a Wavefront OBJ reader, an incremental convex hull, an XML subset
// (elements + attributes: all a URDF needs) and a block-style YAML subset (mappings, sequences, scalars: what PyYAML's
// default dump of scene.yaml / object_library.yaml uses).  Host-side set-up code, double precision.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/jogramop_b200.h"
#include "jg_handles.h"

int jg_fail(int code, const char* fmt, ...);

namespace {

const double HULL_MARGIN = 0.001;   // gUrdfDefaultCollisionMargin

struct P3 { double x, y, z; };
inline P3 operator-(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline P3 cross(P3 a, P3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline bool lex_less(const P3& a, const P3& b) { return a.x != b.x ? a.x < b.x : (a.y != b.y ? a.y < b.y : a.z < b.z); }
inline bool same(const P3& a, const P3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

struct M4 { double m[16]; };   // row-major 4x4
M4 m4_identity() { M4 r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1; return r; }
M4 m4_mul(const M4& a, const M4& b) {
    M4 r{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a.m[4 * i + k] * b.m[4 * k + j];
            r.m[4 * i + j] = s;
        }
    return r;
}
M4 m4_rigid_inverse(const M4& a) {
    M4 r = m4_identity();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[4 * i + j] = a.m[4 * j + i];
    for (int i = 0; i < 3; ++i) r.m[4 * i + 3] = -(r.m[4 * i] * a.m[3] + r.m[4 * i + 1] * a.m[7] + r.m[4 * i + 2] * a.m[11]);
    return r;
}
P3 m4_apply(const M4& T, P3 p) {
    return {T.m[0] * p.x + T.m[1] * p.y + T.m[2] * p.z + T.m[3], T.m[4] * p.x + T.m[5] * p.y + T.m[6] * p.z + T.m[7],
            T.m[8] * p.x + T.m[9] * p.y + T.m[10] * p.z + T.m[11]};
}
// URDF fixed-axis XYZ: R = Rz(yaw) Ry(pitch) Rx(roll)
M4 m4_from_xyz_rpy(const double xyz[3], const double rpy[3]) {
    const double cr = std::cos(rpy[0]), sr = std::sin(rpy[0]), cp = std::cos(rpy[1]), sp = std::sin(rpy[1]),
                 cy = std::cos(rpy[2]), sy = std::sin(rpy[2]);
    M4 T = m4_identity();
    T.m[0] = cy * cp; T.m[1] = cy * sp * sr - sy * cr; T.m[2] = cy * sp * cr + sy * sr;  T.m[3] = xyz[0];
    T.m[4] = sy * cp; T.m[5] = sy * sp * sr + cy * cr; T.m[6] = sy * sp * cr - cy * sr;  T.m[7] = xyz[1];
    T.m[8] = -sp;     T.m[9] = cp * sr;                T.m[10] = cp * cr;                T.m[11] = xyz[2];
    return T;
}

bool read_file(const std::string& path, std::string& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}
std::string dir_of(const std::string& path) {
    const size_t k = path.find_last_of("/\\");
    return k == std::string::npos ? std::string(".") : path.substr(0, k);
}
std::string join_path(const std::string& dir, std::string rel) {
    std::replace(rel.begin(), rel.end(), '\\', '/');      // fix_paths.py: some files carry backslashes
    if (!rel.empty() && rel[0] == '/') return rel;
    return dir + "/" + rel;
}
std::vector<double> parse_doubles(const std::string& s) {
    std::vector<double> v;
    const char* p = s.c_str();
    char* end = nullptr;
    for (;;) {
        const double d = std::strtod(p, &end);
        if (end == p) break;
        v.push_back(d);
        p = end;
    }
    return v;
}

// ------------------------------------------------------------------------------------------------ OBJ
struct ObjShape {
    std::vector<P3> v;                       // the vertices the shape's faces reference, in file order
    std::vector<std::array<int, 3>> f;       // triangles (fan-triangulated polygons), indices into v
};

// v / f (a, a/b, a/b/c, a//c, negative indices, polygons); shapes split on `o` and `g`; everything else ignored
bool read_obj(const std::string& path, std::vector<ObjShape>& shapes, std::string& err) {
    std::string text;
    if (!read_file(path, text)) { err = "cannot open " + path; return false; }
    std::vector<P3> all;
    std::vector<std::vector<std::array<int, 3>>> groups;
    int cur = -1;
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
        size_t i = 0;
        while (i < line.size() && (line[i] == ' ' || line[i] == '\t')) ++i;
        if (i >= line.size() || line[i] == '#') continue;
        const size_t k = line.find_first_of(" \t\r", i);
        const std::string key = line.substr(i, k == std::string::npos ? std::string::npos : k - i);
        const std::string rest = k == std::string::npos ? "" : line.substr(k);
        if (key == "v") {
            const std::vector<double> d = parse_doubles(rest);
            if (d.size() < 3) { err = "bad vertex line in " + path; return false; }
            all.push_back({d[0], d[1], d[2]});
        } else if (key == "o" || key == "g") {
            groups.emplace_back();
            cur = (int)groups.size() - 1;
        } else if (key == "f") {
            if (cur < 0) { groups.emplace_back(); cur = 0; }
            std::vector<int> ids;
            std::istringstream fs(rest);
            std::string tok;
            while (fs >> tok) {
                const int v = std::atoi(tok.c_str());      // stops at the first '/'
                if (v == 0) { err = "bad face index in " + path; return false; }
                ids.push_back(v > 0 ? v - 1 : (int)all.size() + v);
            }
            for (size_t t = 1; t + 1 < ids.size(); ++t) groups[cur].push_back({ids[0], ids[t], ids[t + 1]});
        }
    }
    for (const auto& g : groups) {
        if (g.empty()) continue;
        ObjShape sh;
        std::map<int, int> remap;            // global id -> local id, local ids in ascending global order (np.unique)
        for (const auto& t : g)
            for (int c = 0; c < 3; ++c) {
                if (t[c] < 0 || t[c] >= (int)all.size()) { err = "face references a missing vertex in " + path; return false; }
                remap[t[c]] = 0;
            }
        int n = 0;
        for (auto& kv : remap) { kv.second = n++; sh.v.push_back(all[kv.first]); }
        for (const auto& t : g) sh.f.push_back({remap[t[0]], remap[t[1]], remap[t[2]]});
        shapes.push_back(std::move(sh));
    }
    if (shapes.empty() && !all.empty()) {
        ObjShape sh;
        sh.v = all;
        shapes.push_back(std::move(sh));
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ convex hull
// Vertices of the convex hull of `pts`, in lexicographic order (ingest.convex_hull_vertices: np.unique + qhull).
// Incremental hull with a face list; a point counts as outside a face only beyond eps, so points lying IN a face or
// on an edge are not vertices (like qhull's merged facets).  Support functions -- all the engine needs -- do not
// depend on how such ties are broken.  Flat / tiny inputs: all unique points are returned (support is still exact).
std::vector<P3> hull_vertices(std::vector<P3> pts) {
    std::sort(pts.begin(), pts.end(), lex_less);
    pts.erase(std::unique(pts.begin(), pts.end(), same), pts.end());
    const int n = (int)pts.size();
    if (n <= 4) return pts;
    P3 lo = pts[0], hi = pts[0];
    for (const P3& p : pts) {
        lo = {std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z)};
        hi = {std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z)};
    }
    const double diag = std::sqrt(dot(hi - lo, hi - lo));
    const double eps = 1e-11 * std::max(diag, 1e-30);
    // initial tetrahedron from extreme points
    int i0 = 0, i1 = 0, i2 = -1, i3 = -1;
    double best = -1;
    for (int i = 0; i < n; ++i) { const double d = dot(pts[i] - pts[i0], pts[i] - pts[i0]); if (d > best) { best = d; i1 = i; } }
    best = -1;
    const P3 e01 = pts[i1] - pts[i0];
    for (int i = 0; i < n; ++i) { const P3 c = cross(e01, pts[i] - pts[i0]); const double d = dot(c, c); if (d > best) { best = d; i2 = i; } }
    if (!(best > eps * eps * dot(e01, e01))) return pts;          // collinear
    const P3 nrm = cross(e01, pts[i2] - pts[i0]);
    best = -1;
    for (int i = 0; i < n; ++i) { const double d = std::fabs(dot(nrm, pts[i] - pts[i0])); if (d > best) { best = d; i3 = i; } }
    if (!(best > eps * std::sqrt(dot(nrm, nrm)))) return pts;     // coplanar
    struct Face { int a, b, c; P3 n; double off; bool alive; };
    std::vector<Face> faces;
    auto add_face = [&](int a, int b, int c) {
        P3 nn = cross(pts[b] - pts[a], pts[c] - pts[a]);
        const double l = std::sqrt(dot(nn, nn));
        if (l > 0) nn = {nn.x / l, nn.y / l, nn.z / l};
        faces.push_back({a, b, c, nn, dot(nn, pts[a]), true});
    };
    if (dot(nrm, pts[i3] - pts[i0]) > 0) std::swap(i1, i2);       // make (i0,i1,i2) look away from i3
    add_face(i0, i1, i2); add_face(i0, i3, i1); add_face(i1, i3, i2); add_face(i2, i3, i0);
    std::vector<char> used(n, 0);
    used[i0] = used[i1] = used[i2] = used[i3] = 1;
    for (int p = 0; p < n; ++p) {
        if (used[p]) continue;
        std::vector<int> vis;
        for (int f = 0; f < (int)faces.size(); ++f)
            if (faces[f].alive && dot(faces[f].n, pts[p]) - faces[f].off > eps) vis.push_back(f);
        if (vis.empty()) continue;
        // horizon = directed edges of visible faces whose reverse edge belongs to no visible face
        std::set<std::pair<int, int>> edges;
        for (int f : vis) {
            const Face& F = faces[f];
            edges.insert({F.a, F.b}); edges.insert({F.b, F.c}); edges.insert({F.c, F.a});
        }
        std::vector<std::pair<int, int>> horizon;
        for (const auto& e : edges)
            if (!edges.count({e.second, e.first})) horizon.push_back(e);
        for (int f : vis) faces[f].alive = false;
        for (const auto& e : horizon) add_face(e.first, e.second, p);
        used[p] = 1;
    }
    // A point taken early can end up INSIDE a planar face of the final hull or on one of its edges (its incident
    // triangles are then all coplanar / span two planes): not a vertex (qhull merges such facets).  Extreme points are
    // the ones whose incident face normals span all three dimensions.
    std::vector<std::vector<int>> inc(n);
    for (int f = 0; f < (int)faces.size(); ++f)
        if (faces[f].alive) { inc[faces[f].a].push_back(f); inc[faces[f].b].push_back(f); inc[faces[f].c].push_back(f); }
    std::vector<P3> out;
    for (int i = 0; i < n; ++i) {
        if (inc[i].empty()) continue;
        const double tol = 1e-9;
        P3 n1{0, 0, 0}, n12{0, 0, 0};
        int rank = 0;
        for (int f : inc[i]) {
            const P3 nf = faces[f].n;
            if (!(dot(nf, nf) > 0.5)) continue;                     // degenerate sliver: no normal
            if (rank == 0) { n1 = nf; rank = 1; }
            else if (rank == 1) { const P3 c = cross(n1, nf); if (dot(c, c) > tol * tol) { n12 = c; rank = 2; } }
            else if (std::fabs(dot(n12, nf)) > tol * std::sqrt(dot(n12, n12))) { rank = 3; break; }
        }
        if (rank == 3) out.push_back(pts[i]);
    }
    if (out.size() < 4) return pts;
    return out;
}

// connected components of a triangle mesh after welding coincident vertices (quirk Q1: ungrouped OBJ policy)
std::vector<std::vector<P3>> split_components(const ObjShape& sh, double tol) {
    const int n = (int)sh.v.size();
    std::map<std::array<long long, 3>, int> weld;
    std::vector<int> wid(n);
    for (int i = 0; i < n; ++i) {
        const std::array<long long, 3> key = {(long long)std::llround(sh.v[i].x / tol), (long long)std::llround(sh.v[i].y / tol),
                                              (long long)std::llround(sh.v[i].z / tol)};
        auto it = weld.find(key);
        if (it == weld.end()) it = weld.emplace(key, (int)weld.size()).first;
        wid[i] = it->second;
    }
    std::vector<int> parent(weld.size());
    for (size_t i = 0; i < parent.size(); ++i) parent[i] = (int)i;
    auto find = [&](int a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
    for (const auto& t : sh.f) {
        const int a = find(wid[t[0]]);
        parent[find(wid[t[1]])] = a;
        parent[find(wid[t[2]])] = a;
    }
    std::vector<int> order;                      // roots in order of first appearance
    std::map<int, std::vector<P3>> comp;
    for (const auto& t : sh.f) {
        const int r = find(wid[t[0]]);
        if (!comp.count(r)) order.push_back(r);
        for (int c = 0; c < 3; ++c) comp[r].push_back(sh.v[t[c]]);
    }
    std::vector<std::vector<P3>> out;
    for (int r : order) out.push_back(comp[r]);
    return out;
}

// ------------------------------------------------------------------------------------------------ XML subset
struct XmlNode {
    std::string name;
    std::map<std::string, std::string> attr;
    std::vector<XmlNode> kids;
    const XmlNode* child(const char* n) const {
        for (const XmlNode& k : kids)
            if (k.name == n) return &k;
        return nullptr;
    }
    std::string get(const char* a, const char* def = "") const {
        auto it = attr.find(a);
        return it == attr.end() ? std::string(def) : it->second;
    }
};

bool parse_xml(const std::string& s, XmlNode& root, std::string& err) {
    std::vector<XmlNode*> stack;
    root = XmlNode();
    root.name = "#document";
    stack.push_back(&root);
    size_t i = 0;
    while (i < s.size()) {
        const size_t lt = s.find('<', i);
        if (lt == std::string::npos) break;
        if (s.compare(lt, 4, "<!--") == 0) {
            const size_t e = s.find("-->", lt);
            if (e == std::string::npos) { err = "unterminated comment"; return false; }
            i = e + 3;
            continue;
        }
        if (s.compare(lt, 2, "<?") == 0 || s.compare(lt, 2, "<!") == 0) {
            const size_t e = s.find('>', lt);
            if (e == std::string::npos) { err = "unterminated declaration"; return false; }
            i = e + 1;
            continue;
        }
        const size_t gt = s.find('>', lt);
        if (gt == std::string::npos) { err = "unterminated tag"; return false; }
        std::string tag = s.substr(lt + 1, gt - lt - 1);
        i = gt + 1;
        if (!tag.empty() && tag[0] == '/') {
            if (stack.size() <= 1) { err = "unbalanced closing tag"; return false; }
            stack.pop_back();
            continue;
        }
        bool self_close = false;
        if (!tag.empty() && tag.back() == '/') { self_close = true; tag.pop_back(); }
        XmlNode node;
        size_t p = 0;
        while (p < tag.size() && !isspace((unsigned char)tag[p])) ++p;
        node.name = tag.substr(0, p);
        while (p < tag.size()) {
            while (p < tag.size() && isspace((unsigned char)tag[p])) ++p;
            size_t q = p;
            while (q < tag.size() && tag[q] != '=' && !isspace((unsigned char)tag[q])) ++q;
            if (q == p) break;
            const std::string key = tag.substr(p, q - p);
            while (q < tag.size() && (isspace((unsigned char)tag[q]) || tag[q] == '=')) ++q;
            if (q >= tag.size() || (tag[q] != '"' && tag[q] != '\'')) { err = "attribute without quoted value in <" + node.name + ">"; return false; }
            const char quote = tag[q];
            const size_t e = tag.find(quote, q + 1);
            if (e == std::string::npos) { err = "unterminated attribute value"; return false; }
            node.attr[key] = tag.substr(q + 1, e - q - 1);
            p = e + 1;
        }
        stack.back()->kids.push_back(std::move(node));
        if (!self_close) stack.push_back(&stack.back()->kids.back());
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ YAML subset
struct YNode {
    enum Kind { NUL, SCALAR, MAP, SEQ } kind = NUL;
    std::string s;
    std::vector<std::pair<std::string, YNode>> map;
    std::vector<YNode> seq;
    const YNode* get(const char* key) const {
        for (const auto& kv : map)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct YLine { int indent; std::string text; };

std::string y_unquote(std::string v) {
    while (!v.empty() && (v.back() == ' ' || v.back() == '\r')) v.pop_back();
    if (v.size() >= 2 && ((v.front() == '\'' && v.back() == '\'') || (v.front() == '"' && v.back() == '"'))) v = v.substr(1, v.size() - 2);
    return v;
}

// block-style subset: `key: value`, `key:` + nested block (a sequence may sit at the key's own indent, PyYAML style),
// `- item` where the item continues on the same line (`- - 1.0`, `- key: value`)
struct YParser {
    std::vector<YLine> lines;
    size_t pos = 0;
    std::string err;
    bool parse(YNode& out, int indent) {
        if (pos >= lines.size()) return true;
        YLine& L = lines[pos];
        if (L.indent != indent) { err = "unexpected indentation"; return false; }
        if (L.text == "-" || L.text.compare(0, 2, "- ") == 0) {
            out.kind = YNode::SEQ;
            while (pos < lines.size() && lines[pos].indent == indent &&
                   (lines[pos].text == "-" || lines[pos].text.compare(0, 2, "- ") == 0)) {
                YLine& I = lines[pos];
                YNode item;
                if (I.text == "-") {
                    ++pos;
                    if (pos < lines.size() && lines[pos].indent > indent && !parse(item, lines[pos].indent)) return false;
                } else {
                    size_t k = 2;
                    while (k < I.text.size() && I.text[k] == ' ') ++k;
                    I.indent = indent + (int)k;           // the item continues on this line, two (or more) columns in
                    I.text = I.text.substr(k);
                    if (!parse(item, I.indent)) return false;
                }
                out.seq.push_back(std::move(item));
            }
            return true;
        }
        const size_t colon = L.text.find(':');
        const bool is_key = colon != std::string::npos && (colon + 1 == L.text.size() || L.text[colon + 1] == ' ') &&
                            L.text[0] != '\'' && L.text[0] != '"';
        if (!is_key) {
            out.kind = YNode::SCALAR;
            out.s = y_unquote(L.text);
            if (out.s == "null" || out.s == "~") out.kind = YNode::NUL;
            ++pos;
            return true;
        }
        out.kind = YNode::MAP;
        while (pos < lines.size() && lines[pos].indent == indent) {
            const std::string t = lines[pos].text;
            if (t == "-" || t.compare(0, 2, "- ") == 0) break;     // the enclosing sequence goes on
            const size_t c = t.find(':');
            if (c == std::string::npos) { err = "expected `key:` in a mapping, got: " + t; return false; }
            const std::string key = y_unquote(t.substr(0, c));
            std::string rest = c + 1 < t.size() ? t.substr(c + 1) : "";
            size_t k = 0;
            while (k < rest.size() && rest[k] == ' ') ++k;
            rest = rest.substr(k);
            YNode val;
            ++pos;
            if (!rest.empty()) {
                val.kind = YNode::SCALAR;
                val.s = y_unquote(rest);
                if (val.s == "null" || val.s == "~") val.kind = YNode::NUL;
            } else if (pos < lines.size()) {
                const YLine& N = lines[pos];
                const bool dash = N.text == "-" || N.text.compare(0, 2, "- ") == 0;
                if (N.indent > indent || (N.indent == indent && dash)) {
                    if (!parse(val, N.indent)) return false;
                }
            }
            out.map.emplace_back(key, std::move(val));
        }
        return true;
    }
};

bool parse_yaml(const std::string& text, YNode& root, std::string& err) {
    YParser P;
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
        size_t i = 0;
        while (i < line.size() && line[i] == ' ') ++i;
        if (i >= line.size() || line[i] == '#') continue;
        if (line.compare(i, 3, "---") == 0 || line.compare(i, 3, "...") == 0) continue;
        P.lines.push_back({(int)i, line.substr(i)});
    }
    if (P.lines.empty()) { root = YNode(); return true; }
    if (!P.parse(root, P.lines[0].indent)) { err = P.err; return false; }
    if (P.pos != P.lines.size()) { err = "trailing content at line: " + P.lines[P.pos].text; return false; }
    return true;
}

bool y_pose(const YNode* n, M4& T) {
    if (!n || n->kind != YNode::SEQ || n->seq.size() != 4) return false;
    for (int r = 0; r < 4; ++r) {
        const YNode& row = n->seq[r];
        if (row.kind != YNode::SEQ || row.seq.size() != 4) return false;
        for (int c = 0; c < 4; ++c) {
            if (row.seq[c].kind != YNode::SCALAR) return false;
            T.m[4 * r + c] = std::strtod(row.seq[c].s.c_str(), nullptr);
        }
    }
    return true;
}

M4 default_robot_pose() {   // scenario.py:69-80: Rz(+90 deg), t = (1, 0, 0.05)
    M4 T = m4_identity();
    T.m[0] = 0; T.m[1] = -1; T.m[3] = 1.0;
    T.m[4] = 1; T.m[5] = 0;  T.m[7] = 0.0;
    T.m[11] = 0.05;
    return T;
}

struct SceneBody { std::string type, vhacd_fn, mesh_fn; M4 pose; bool target; };

bool load_scene_bodies(const char* scene_yaml, std::vector<SceneBody>& bodies, std::string& err) {
    std::string text;
    if (!read_file(scene_yaml, text)) { err = std::string("cannot open ") + scene_yaml; return false; }
    YNode scene;
    if (!parse_yaml(text, scene, err)) { err = std::string(scene_yaml) + ": " + err; return false; }
    const YNode* libfn = scene.get("object_library_fn");
    if (!libfn || libfn->kind != YNode::SCALAR) { err = "scene.yaml has no object_library_fn"; return false; }
    const std::string lib_path = join_path(dir_of(scene_yaml), libfn->s);
    if (!read_file(lib_path, text)) { err = "cannot open " + lib_path; return false; }
    YNode lib;
    if (!parse_yaml(text, lib, err)) { err = lib_path + ": " + err; return false; }
    const YNode* lobjs = lib.get("objects");
    if (!lobjs || lobjs->kind != YNode::SEQ) { err = lib_path + ": no `objects` list"; return false; }
    const std::string lib_dir = dir_of(lib_path);
    std::map<std::string, const YNode*> by_id;
    for (const YNode& o : lobjs->seq) {
        const YNode* id = o.get("identifier");
        if (id && id->kind == YNode::SCALAR) by_id[id->s] = &o;
    }
    // body order of create_cpp_export.py:41: objects first, then bg_objects
    for (int pass = 0; pass < 2; ++pass) {
        const YNode* list = scene.get(pass == 0 ? "objects" : "bg_objects");
        if (!list || list->kind != YNode::SEQ) continue;
        for (const YNode& e : list->seq) {
            const YNode* t = e.get("object_type");
            SceneBody b;
            if (!t || t->kind != YNode::SCALAR || !y_pose(e.get("pose"), b.pose)) { err = "scene entry without object_type / 4x4 pose"; return false; }
            b.type = t->s;
            b.target = pass == 0;
            auto it = by_id.find(b.type);
            if (it == by_id.end()) { err = "object type " + b.type + " not in " + lib_path; return false; }
            const YNode* sc = it->second->get("scale");
            if (sc && sc->kind == YNode::SCALAR && std::strtod(sc->s.c_str(), nullptr) != 1.0) {
                err = "did not implement scale feature for VHACD export";      // create_cpp_export.py:33-34
                return false;
            }
            const YNode* vf = it->second->get("vhacd_fn");
            const YNode* mf = it->second->get("mesh_fn");
            if (vf && vf->kind == YNode::SCALAR) b.vhacd_fn = join_path(lib_dir, vf->s);
            if (mf && mf->kind == YNode::SCALAR) b.mesh_fn = join_path(lib_dir, mf->s);
            bodies.push_back(std::move(b));
        }
    }
    return true;
}

}  // namespace

// ================================================================================================ C-ABI
extern "C" jg_robot* jg_robot_load(const char* urdf_path, const char* mesh_root) {
    if (!urdf_path) { jg_fail(JG_ERR_INVALID, "jg_robot_load: urdf_path is NULL"); return nullptr; }
    std::string text, err;
    if (!read_file(urdf_path, text)) { jg_fail(JG_ERR_INVALID, "jg_robot_load: cannot open %s", urdf_path); return nullptr; }
    XmlNode doc;
    if (!parse_xml(text, doc, err)) { jg_fail(JG_ERR_INVALID, "jg_robot_load: %s: %s", urdf_path, err.c_str()); return nullptr; }
    const XmlNode* robot = doc.child("robot");
    if (!robot) { jg_fail(JG_ERR_INVALID, "jg_robot_load: %s has no <robot> element", urdf_path); return nullptr; }
    const std::string base_dir = mesh_root ? std::string(mesh_root) : dir_of(urdf_path);
    std::map<std::string, const XmlNode*> links;
    std::vector<const XmlNode*> joints;
    std::set<std::string> children;
    for (const XmlNode& k : robot->kids) {
        if (k.name == "link") links[k.get("name")] = &k;
        if (k.name == "joint") {
            joints.push_back(&k);
            if (k.child("child")) children.insert(k.child("child")->get("link"));
        }
    }
    std::string base;
    int n_roots = 0;
    for (const XmlNode& k : robot->kids)       // file order
        if (k.name == "link" && !children.count(k.get("name"))) { if (n_roots++ == 0) base = k.get("name"); }
    if (n_roots != 1) { jg_fail(JG_ERR_INVALID, "jg_robot_load: expected a single root link, found %d", n_roots); return nullptr; }
    // pyBullet numbering: depth first from the base, children in file order
    std::vector<const XmlNode*> order;
    std::vector<std::string> stack_links = {base};
    {
        struct Rec { static void visit(const std::string& link, const std::vector<const XmlNode*>& joints, std::vector<const XmlNode*>& order) {
            for (const XmlNode* j : joints)
                if (j->child("parent") && j->child("parent")->get("link") == link) {
                    order.push_back(j);
                    visit(j->child("child")->get("link"), joints, order);
                }
        } };
        Rec::visit(base, joints, order);
    }
    const int L = (int)order.size();
    if (L <= 0) { jg_fail(JG_ERR_INVALID, "jg_robot_load: no joints"); return nullptr; }
    std::map<std::string, int> link_index;
    link_index[base] = -1;
    for (int i = 0; i < L; ++i) link_index[order[i]->child("child")->get("link")] = i;
    jg_robot tmp;
    std::vector<int> jt(L), jp(L), qidx(L, -1);
    std::vector<double> jo(12 * L), ja(3 * L, 0.0), lim;
    int dof = 0;
    bool with_platform = false;
    for (int i = 0; i < L; ++i) {
        const XmlNode* j = order[i];
        const std::string type = j->get("type");
        jt[i] = type == "fixed" ? JG_JOINT_FIXED : ((type == "revolute" || type == "continuous") ? JG_JOINT_REVOLUTE
                : (type == "prismatic" ? JG_JOINT_PRISMATIC : -1));
        if (jt[i] < 0) { jg_fail(JG_ERR_UNSUPPORTED, "jg_robot_load: joint type '%s' is not supported", type.c_str()); return nullptr; }
        jp[i] = link_index[j->child("parent")->get("link")];
        double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0};
        if (const XmlNode* o = j->child("origin")) {
            const std::vector<double> a = parse_doubles(o->get("xyz", "0 0 0")), b = parse_doubles(o->get("rpy", "0 0 0"));
            for (int k = 0; k < 3 && k < (int)a.size(); ++k) xyz[k] = a[k];
            for (int k = 0; k < 3 && k < (int)b.size(); ++k) rpy[k] = b[k];
        }
        const M4 T = m4_from_xyz_rpy(xyz, rpy);
        for (int k = 0; k < 12; ++k) jo[12 * i + k] = T.m[k];
        if (jt[i] != JG_JOINT_FIXED) {
            double ax[3] = {1, 0, 0};
            if (const XmlNode* a = j->child("axis")) {
                const std::vector<double> v = parse_doubles(a->get("xyz", "1 0 0"));
                for (int k = 0; k < 3 && k < (int)v.size(); ++k) ax[k] = v[k];
            }
            for (int k = 0; k < 3; ++k) ja[3 * i + k] = ax[k];
        }
        tmp.joint_names.push_back(j->get("name"));
        tmp.link_names.push_back(j->child("child")->get("link"));
        if (tmp.link_names.back() == "mobile_platform") with_platform = true;
    }
    // configuration vector = the movable joints in link order except the gripper fingers (simulation.py:156-164:
    // arm_joint_ids [0,1,3..9] / [0..6]); finger joints follow params.finger_open
    for (int i = 0; i < L; ++i) {
        if (jt[i] == JG_JOINT_FIXED) continue;
        if (tmp.joint_names[i].find("finger") != std::string::npos) { qidx[i] = -2; continue; }
        qidx[i] = dof++;
        double lo = 0, hi = 0;
        if (const XmlNode* l = order[i]->child("limit")) {
            lo = std::strtod(l->get("lower", "0").c_str(), nullptr);
            hi = std::strtod(l->get("upper", "0").c_str(), nullptr);
        }
        lim.push_back(lo);
        lim.push_back(hi);
    }
    if (dof <= 0) { jg_fail(JG_ERR_INVALID, "jg_robot_load: the robot has no movable arm joint"); return nullptr; }
    // collision shapes, link -1 (base) first
    std::vector<int> s_link, s_off = {0}, p_off = {0};
    std::vector<double> s_margin, s_pm, core, plane;
    for (int i = -1; i < L; ++i) {
        const XmlNode* link = links[i < 0 ? base : tmp.link_names[i]];
        if (!link) continue;
        for (const XmlNode& col : link->kids) {
            if (col.name != "collision") continue;
            double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0};
            if (const XmlNode* o = col.child("origin")) {
                const std::vector<double> a = parse_doubles(o->get("xyz", "0 0 0")), b = parse_doubles(o->get("rpy", "0 0 0"));
                for (int k = 0; k < 3 && k < (int)a.size(); ++k) xyz[k] = a[k];
                for (int k = 0; k < 3 && k < (int)b.size(); ++k) rpy[k] = b[k];
            }
            const M4 T = m4_from_xyz_rpy(xyz, rpy);
            const XmlNode* geom = col.child("geometry");
            const XmlNode* mesh = geom ? geom->child("mesh") : nullptr;
            const XmlNode* box = geom ? geom->child("box") : nullptr;
            if (mesh) {
                std::vector<double> sc = parse_doubles(mesh->get("scale", "1 1 1"));
                if (sc.size() < 3) sc = {1, 1, 1};
                std::vector<ObjShape> shapes;
                if (!read_obj(join_path(base_dir, mesh->get("filename")), shapes, err)) { jg_fail(JG_ERR_INVALID, "jg_robot_load: %s", err.c_str()); return nullptr; }
                for (ObjShape& sh : shapes) {
                    for (P3& p : sh.v) p = {p.x * sc[0], p.y * sc[1], p.z * sc[2]};
                    const std::vector<P3> h = hull_vertices(sh.v);
                    s_link.push_back(i); s_margin.push_back(HULL_MARGIN); s_pm.push_back(HULL_MARGIN);
                    for (const P3& p0 : h) {
                        const P3 p = m4_apply(T, p0);
                        core.insert(core.end(), {p.x, p.y, p.z});
                        plane.insert(plane.end(), {p.x, p.y, p.z});
                    }
                    s_off.push_back((int)core.size() / 3);
                    p_off.push_back((int)plane.size() / 3);
                }
            } else if (box) {
                const std::vector<double> sz = parse_doubles(box->get("size", "0 0 0"));
                if (sz.size() < 3) { jg_fail(JG_ERR_INVALID, "jg_robot_load: <box> without size"); return nullptr; }
                s_link.push_back(i); s_margin.push_back(HULL_MARGIN); s_pm.push_back(0.0);
                for (int sx = -1; sx <= 1; sx += 2)
                    for (int sy = -1; sy <= 1; sy += 2)
                        for (int sz_ = -1; sz_ <= 1; sz_ += 2) {
                            const P3 c = m4_apply(T, {sx * (0.5 * sz[0] - HULL_MARGIN), sy * (0.5 * sz[1] - HULL_MARGIN), sz_ * (0.5 * sz[2] - HULL_MARGIN)});
                            const P3 e = m4_apply(T, {sx * 0.5 * sz[0], sy * 0.5 * sz[1], sz_ * 0.5 * sz[2]});
                            core.insert(core.end(), {c.x, c.y, c.z});
                            plane.insert(plane.end(), {e.x, e.y, e.z});
                        }
                s_off.push_back((int)core.size() / 3);
                p_off.push_back((int)plane.size() / 3);
            } else {
                jg_fail(JG_ERR_UNSUPPORTED, "jg_robot_load: only <mesh> and <box> collision geometry is supported");
                return nullptr;
            }
        }
    }
    // FrankaRobot.in_self_collision's ordered link pairs (simulation.py:439-464)
    std::vector<int> pairs;
    int ee = -1;
    {
        const int max_link = with_platform ? 14 : 11;
        const std::vector<int> ignore = with_platform ? std::vector<int>{0, 1, 10, 14} : std::vector<int>{7, 11};
        const std::vector<int> first = with_platform ? std::vector<int>{2, 3, 4, 5, 6, 7, 8} : std::vector<int>{1, 2, 3, 4, 5};
        if (L > max_link) {
            for (int f : first)
                for (int c = f + 2; c <= max_link; ++c)
                    if (std::find(ignore.begin(), ignore.end(), c) == ignore.end()) { pairs.push_back(f); pairs.push_back(c); }
            ee = max_link;
        }
    }
    const int S = (int)s_link.size();
    jg_robot* r = jg_robot_create(L, jt.data(), jp.data(), jo.data(), ja.data(), qidx.data(), lim.data(), dof, S, s_link.data(),
                                  s_margin.data(), s_off.data(), core.data(), p_off.data(), plane.data(), s_pm.data(),
                                  (int)pairs.size() / 2, pairs.data());
    if (!r) return nullptr;
    r->link_names = tmp.link_names;
    r->joint_names = tmp.joint_names;
    r->ee_link = ee;
    return r;
}

static jg_scene* scene_load_impl(const char* scene_yaml, const double* robot_pose, int policy, bool mesh, int mesh_source, double mesh_margin) {
    if (!scene_yaml) { jg_fail(JG_ERR_INVALID, "jg_scene_load: scene_yaml is NULL"); return nullptr; }
    std::vector<SceneBody> bodies;
    std::string err;
    if (!load_scene_bodies(scene_yaml, bodies, err)) {
        jg_fail(err.find("scale feature") != std::string::npos ? JG_ERR_UNSUPPORTED : JG_ERR_INVALID, "jg_scene_load: %s", err.c_str());
        return nullptr;
    }
    M4 rp = default_robot_pose();
    if (robot_pose) memcpy(rp.m, robot_pose, sizeof(double) * 16);
    const M4 to_robot = m4_rigid_inverse(rp);
    const double plane[4] = {rp.m[8], rp.m[9], rp.m[10], rp.m[11]};   // world z = 0 in the robot frame
    std::vector<int> off = {0}, piece_body, tris;
    std::vector<double> verts, margin, tri_verts;
    std::vector<std::string> names;
    int n_targets = 0;
    for (size_t b = 0; b < bodies.size(); ++b) {
        const SceneBody& B = bodies[b];
        names.push_back(B.type);
        if (B.target) ++n_targets;
        const M4 T = m4_mul(to_robot, B.pose);
        std::vector<ObjShape> shapes;
        const std::string& fn = (mesh && mesh_source == JG_MESH_VISUAL) ? B.mesh_fn : B.vhacd_fn;
        if (fn.empty() || !read_obj(fn, shapes, err)) { jg_fail(JG_ERR_INVALID, "jg_scene_load: body %s: %s", B.type.c_str(), fn.empty() ? "no mesh file in the library" : err.c_str()); return nullptr; }
        if (mesh) {
            for (const ObjShape& sh : shapes) {
                const int base = (int)tri_verts.size() / 3;
                for (const P3& p0 : sh.v) { const P3 p = m4_apply(T, p0); tri_verts.insert(tri_verts.end(), {p.x, p.y, p.z}); }
                for (const auto& t : sh.f) tris.insert(tris.end(), {base + t[0], base + t[1], base + t[2]});
            }
            continue;
        }
        std::vector<std::vector<P3>> sets;
        if (shapes.size() == 1 && policy == JG_OBJ_CONNECTED_COMPONENTS) sets = split_components(shapes[0], 1e-6);
        else for (const ObjShape& sh : shapes) sets.push_back(sh.v);
        for (const auto& set : sets) {
            const std::vector<P3> h = hull_vertices(set);
            for (const P3& p0 : h) { const P3 p = m4_apply(T, p0); verts.insert(verts.end(), {p.x, p.y, p.z}); }
            off.push_back((int)verts.size() / 3);
            margin.push_back(HULL_MARGIN);
            piece_body.push_back((int)b);
        }
    }
    jg_scene* s = mesh ? jg_scene_create_mesh((int)tri_verts.size() / 3, tri_verts.data(), (int)tris.size() / 3, tris.data(), mesh_margin, plane)
                       : jg_scene_create((int)margin.size(), off.data(), verts.data(), margin.data(), plane);
    if (!s) return nullptr;
    s->piece_body = piece_body;
    s->body_names = names;
    s->n_target_bodies = n_targets;
    return s;
}

extern "C" jg_scene* jg_scene_load(const char* scene_yaml, const double robot_pose[16]) {
    return scene_load_impl(scene_yaml, robot_pose, JG_OBJ_SINGLE_HULL, false, 0, 0.0);
}
extern "C" jg_scene* jg_scene_load_ex(const char* scene_yaml, const double robot_pose[16], int ungrouped_obj_policy) {
    if (ungrouped_obj_policy != JG_OBJ_SINGLE_HULL && ungrouped_obj_policy != JG_OBJ_CONNECTED_COMPONENTS) {
        jg_fail(JG_ERR_INVALID, "jg_scene_load_ex: unknown ungrouped_obj_policy %d", ungrouped_obj_policy);
        return nullptr;
    }
    return scene_load_impl(scene_yaml, robot_pose, ungrouped_obj_policy, false, 0, 0.0);
}
extern "C" jg_scene* jg_scene_load_mesh(const char* scene_yaml, const double robot_pose[16], int mesh_source, double margin) {
    if (mesh_source != JG_MESH_VHACD && mesh_source != JG_MESH_VISUAL) {
        jg_fail(JG_ERR_INVALID, "jg_scene_load_mesh: unknown mesh_source %d", mesh_source);
        return nullptr;
    }
    return scene_load_impl(scene_yaml, robot_pose, JG_OBJ_SINGLE_HULL, true, mesh_source, margin);
}

extern "C" int jg_robot_describe(const jg_robot* r, jg_robot_desc* d) {
    if (!r || !d) return jg_fail(JG_ERR_INVALID, "jg_robot_describe: NULL argument");
    d->n_links = r->L; d->n_shapes = r->S; d->dof = r->dof; d->n_self_pairs = r->K; d->ee_link = r->ee_link;
    d->joint_type = r->joint_type.data(); d->joint_parent = r->joint_parent.data(); d->joint_qidx = r->joint_qidx.data();
    d->joint_origin = r->joint_origin.data(); d->joint_axis = r->joint_axis.data(); d->q_limits = r->q_limits.data();
    d->shape_link = r->shape_link.data(); d->shape_margin = r->shape_margin.data(); d->shape_vert_offset = r->shape_off.data();
    d->core_verts = r->core.data(); d->shape_plane_offset = r->plane_off.data(); d->plane_verts = r->plane_v.data();
    d->shape_plane_margin = r->plane_margin.data(); d->self_pairs = r->self_pairs.data();
    return JG_OK;
}
extern "C" int jg_robot_link_index(const jg_robot* r, const char* link_name) {
    if (!r || !link_name) return jg_fail(JG_ERR_INVALID, "jg_robot_link_index: NULL argument");
    for (size_t i = 0; i < r->link_names.size(); ++i)
        if (r->link_names[i] == link_name) return (int)i;
    return jg_fail(JG_ERR_INVALID, "jg_robot_link_index: no link named '%s' (names are only known for robots made by jg_robot_load)", link_name);
}
extern "C" int jg_scene_describe(const jg_scene* s, jg_scene_desc* d) {
    if (!s || !d) return jg_fail(JG_ERR_INVALID, "jg_scene_describe: NULL argument");
    d->n_pieces = s->P; d->piece_vert_offset = s->off.data(); d->piece_verts = s->v.data(); d->piece_margin = s->margin.data();
    d->piece_body = s->piece_body.empty() ? nullptr : s->piece_body.data();
    d->n_bodies = (int)s->body_names.size(); d->n_target_bodies = s->n_target_bodies;
    d->has_plane = s->has_plane ? 1 : 0;
    for (int k = 0; k < 4; ++k) d->plane[k] = s->plane[k];
    d->is_mesh = s->is_mesh ? 1 : 0;
    d->n_mesh_verts = (int)s->mesh_v.size() / 3; d->n_mesh_tris = (int)s->mesh_t.size() / 3;
    d->mesh_verts = s->mesh_v.empty() ? nullptr : s->mesh_v.data();
    d->mesh_tris = s->mesh_t.empty() ? nullptr : s->mesh_t.data();
    return JG_OK;
}
