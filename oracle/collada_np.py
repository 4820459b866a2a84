"""ORACLE (test infrastructure only -- never imported by the product path).

numpy/ElementTree restatement of the reference's mesh import
(`include/mpl/demo/load_mesh.hpp:232-293` MeshLoad::load, `:50-89` visitTriangles,
`:30-47` visitVertices) on top of the assimp Collada importer conventions that the
reference relies on ([EXT] assimp, not vendored in /root/reference, version unpinned --
`CMakeLists.txt:25`):

  * one aiMesh per (instance_geometry x primitive group); `<lines>` groups become meshes
    whose faces have 2 indices and are skipped by `load_mesh.hpp:65-66`;
  * node transforms are the nested `<matrix>` elements (row-major 4x4, float32),
    accumulated root->leaf, cast float->double before multiplying (`load_mesh.hpp:60`);
  * `<up_axis>Z_UP` puts the rotation (x,y,z)->(x,z,-y) on the root node; the reference
    can overwrite the root transform with identity (`load_mesh.hpp:256-257`);
  * `<unit meter=..>` scales the root transform (assimp >= 5.0); every file the scripts
    use has either no <unit> or meter="1".
  * shiftToCenter subtracts the mean of all vertices of all meshes
    (`load_mesh.hpp:263-273`); vertices are counted after aiProcess_JoinIdenticalVertices,
    restated here as a per-primitive-group de-duplication on the (position, normal,
    texcoord) VALUE tuple.  parity unpinned: assimp itself is absent, see DESIGN.md.
"""
import xml.etree.ElementTree as ET
import numpy as np

NS = "{http://www.collada.org/2005/11/COLLADASchema}"


def _floats(text, dtype=np.float32):
    return np.array(text.split(), dtype=dtype)


class _Geom:
    def __init__(self, el):
        mesh = el.find(NS + "mesh")
        self.sources = {}
        for src in mesh.findall(NS + "source"):
            fa = src.find(NS + "float_array")
            acc = src.find(NS + "technique_common/" + NS + "accessor")
            stride = int(acc.get("stride", "1"))
            self.sources["#" + src.get("id")] = _floats(fa.text or "").reshape(-1, stride)
        self.vertices = {}
        for v in mesh.findall(NS + "vertices"):
            self.vertices["#" + v.get("id")] = {
                i.get("semantic"): i.get("source") for i in v.findall(NS + "input")}
        self.groups = []  # (kind, per-corner attribute index table, inputs)
        for child in mesh:
            kind = child.tag[len(NS):]
            if kind not in ("triangles", "lines", "polylist", "polygons"):
                continue
            inputs = [(i.get("semantic"), i.get("source"), int(i.get("offset", "0")))
                      for i in child.findall(NS + "input")]
            nstride = max(o for _, _, o in inputs) + 1
            p = child.find(NS + "p")
            idx = np.array((p.text or "").split(), dtype=np.int64).reshape(-1, nstride)
            if kind == "polylist":
                vcount = np.array(child.find(NS + "vcount").text.split(), dtype=np.int64)
            elif kind == "triangles":
                vcount = np.full(len(idx) // 3, 3, dtype=np.int64)
            elif kind == "lines":
                vcount = np.full(len(idx) // 2, 2, dtype=np.int64)
            else:
                raise NotImplementedError(kind)
            self.groups.append((kind, inputs, idx, vcount))

    def corner_attributes(self, group):
        """-> positions [ncorner,3] float32, key [ncorner,k] float32 (pos+normal+uv)."""
        kind, inputs, idx, vcount = group
        cols = []
        pos = None
        for sem, src, off in inputs:
            if sem == "VERTEX":
                for vsem, vsrc in self.vertices[src].items():
                    arr = self.sources[vsrc][idx[:, off]]
                    if vsem == "POSITION":
                        pos = arr[:, :3]
                    cols.append(arr)
            elif sem in ("NORMAL", "TEXCOORD"):
                cols.append(self.sources[src][idx[:, off]])
        return pos, np.concatenate(cols, axis=1)


def _node_matrix(node):
    m = np.eye(4)
    for child in node:
        tag = child.tag[len(NS):]
        if tag == "matrix":
            m = m @ _floats(child.text).reshape(4, 4).astype(np.float64)
        elif tag in ("translate", "rotate", "scale", "lookat", "skew"):
            raise NotImplementedError("COLLADA <%s> transform" % tag)
    return m


def load_collada(path, shift_to_center=False, identity_root_transform=False,
                 return_center=False):
    """-> float64 triangles [n,3,3], exactly the soup `visitTriangles` would emit."""
    root = ET.parse(path).getroot()
    geoms = {"#" + g.get("id"): _Geom(g)
             for g in root.findall(NS + "library_geometries/" + NS + "geometry")
             if g.find(NS + "mesh") is not None}
    up = root.find(NS + "asset/" + NS + "up_axis")
    unit = root.find(NS + "asset/" + NS + "unit")
    root_tf = np.eye(4)
    if not identity_root_transform:
        if up is not None and up.text.strip() == "Z_UP":
            root_tf = root_tf @ np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1.0]])
        elif up is not None and up.text.strip() == "X_UP":
            root_tf = root_tf @ np.array([[0, -1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
        if unit is not None:
            s = float(np.float32(unit.get("meter", "1")))
            root_tf = root_tf @ np.diag([s, s, s, 1.0])
    scene_url = root.find(NS + "scene/" + NS + "instance_visual_scene").get("url")
    vscene = [v for v in root.findall(NS + "library_visual_scenes/" + NS + "visual_scene")
              if "#" + v.get("id") == scene_url][0]

    tri_chunks, vert_chunks = [], []

    def visit(node, tf):
        tf = tf @ _node_matrix(node)
        for inst in node.findall(NS + "instance_geometry"):
            g = geoms.get(inst.get("url"))
            if g is None:
                continue
            for group in g.groups:
                pos, key = g.corner_attributes(group)
                kind, _, _, vcount = group
                # JoinIdenticalVertices: unique (pos, normal, uv) tuples per mesh
                uniq = np.unique(key, axis=0, return_index=True)[1]
                p64 = pos.astype(np.float64)
                wpos = p64 @ tf[:3, :3].T + tf[:3, 3]
                vert_chunks.append(wpos[np.sort(uniq)])
                start = 0
                for n in vcount:
                    if n >= 3:
                        for k in range(2, n):  # fan around corner 0 (load_mesh.hpp:75-81)
                            tri_chunks.append(wpos[[start, start + k - 1, start + k]])
                    start += n
        for child in node.findall(NS + "node"):
            visit(child, tf)

    for node in vscene.findall(NS + "node"):
        visit(node, root_tf)

    tris = np.array(tri_chunks, dtype=np.float64).reshape(-1, 3, 3)
    center = np.zeros(3)
    if shift_to_center:
        allv = np.concatenate(vert_chunks, axis=0)
        center = allv.sum(axis=0) / len(allv)
        tris = tris - center
    if return_center:
        return tris, center
    return tris
