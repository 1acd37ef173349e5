"""Generate the golden vectors under tests/golden/ from the REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's ``autodriver_pointcloud_preprocessor/utils.py`` *verbatim* from
``/root/reference`` under two stub ROS modules (``sensor_msgs.msg``,
``sensor_msgs_py.point_cloud2``) and a ~30-line duck-typed Open3D tensor/point-cloud shim
(the module's own torch/open3d imports are guarded, ``utils.py:9-26``), feeds it seeded
inputs and stores inputs + outputs as small ``.npz`` / ``.json`` fixtures.  These pin the
oracle (and through it the CUDA path) for every stage whose arithmetic lives in the
reference: crop (numpy + torch back ends), dedup (numpy + torch back ends), structured
array -> SoA conversion, vendor field mapping, packed-field layout, rgb helpers.  It also
parses ``pointcloud_preprocessor.py`` with ``ast`` (that file cannot be imported - rclpy /
open3d / cv_bridge are absent) and dumps the declared parameter table.

Nothing from the reference is copied into the repo: only its *outputs* are stored.
"""
from __future__ import annotations

import ast
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def install_stubs():
    """Stub modules so that ``utils.py:6-7`` imports succeed."""
    class PointField:
        INT8, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 = 1, 2, 3, 4, 5, 6, 7, 8

        def __init__(self):
            self.name, self.offset, self.datatype, self.count = "", 0, 0, 1

    class PointCloud2:
        pass

    sm = types.ModuleType("sensor_msgs")
    smm = types.ModuleType("sensor_msgs.msg")
    smm.PointField, smm.PointCloud2 = PointField, PointCloud2
    sm.msg = smm
    smp = types.ModuleType("sensor_msgs_py")
    pc2 = types.ModuleType("sensor_msgs_py.point_cloud2")
    smp.point_cloud2 = pc2
    sys.modules.update({"sensor_msgs": sm, "sensor_msgs.msg": smm, "sensor_msgs_py": smp,
                        "sensor_msgs_py.point_cloud2": pc2})


class ShimTensor:
    """Duck-typed ``o3c.Tensor`` over a numpy array (only what utils.py touches)."""

    def __init__(self, a):
        self.a = a

    is_cpu = True

    def cpu(self):
        return self

    def numpy(self):
        return self.a

    def to(self, *_a, **_k):
        return self

    def to_dlpack(self):
        import torch
        return torch.utils.dlpack.to_dlpack(torch.from_numpy(self.a))

    @staticmethod
    def from_numpy(a):
        return ShimTensor(a)

    @staticmethod
    def from_dlpack(cap):
        import torch
        return ShimTensor(torch.utils.dlpack.from_dlpack(cap).numpy())


class ShimPoint:
    def __init__(self, positions):
        self.positions = ShimTensor(positions)


class ShimCloud:
    device = "CPU:0"

    def __init__(self, positions):
        self.point = ShimPoint(positions)
        self.selected_mask = None
        self.selected_index = None

    def select_by_mask(self, mask):
        self.selected_mask = np.asarray(mask.a).astype(bool)
        return self

    def select_by_index(self, idx, invert=False):
        self.selected_index = np.asarray(idx.a)
        return self


def load_reference_utils():
    install_stubs()
    sys.path.insert(0, REF)
    import importlib
    utils = importlib.import_module("autodriver_pointcloud_preprocessor.utils")
    shim = types.SimpleNamespace(Tensor=ShimTensor, Dtype=types.SimpleNamespace(Bool="bool"))
    utils.o3c = shim
    return utils


def adversarial_points(rng, n=4000):
    p = rng.uniform(-80, 80, size=(n, 3)).astype(np.float32)
    p[:, 2] = rng.uniform(-30, 30, size=n).astype(np.float32)
    # exact boundary values, values one ulp either side, NaN / inf rows, signed zeros
    lo = np.array([-60.0, -60.0, -20.0], dtype=np.float32)
    hi = np.array([60.0, 60.0, 20.0], dtype=np.float32)
    for k in range(60):
        i = rng.integers(0, n)
        ax = k % 3
        base = (lo if (k // 3) % 2 == 0 else hi)[ax]
        v = [base, np.nextafter(base, np.float32(np.inf)), np.nextafter(base, np.float32(-np.inf))][k % 3]
        p[i, ax] = v
    for k in range(30):
        p[rng.integers(0, n), rng.integers(0, 3)] = [np.nan, np.inf, -np.inf][k % 3]
    p[rng.integers(0, n, size=10)] = np.float32(0.0)
    p[rng.integers(0, n, size=10)] = np.float32(-0.0)
    return p


def gen_crop(utils, out):
    rng = np.random.default_rng(1234)
    p = adversarial_points(rng)
    cases = {
        "roi": ([-60.0, -60.0, -20.0], [60.0, 60.0, 20.0]),
        # bounds that are NOT float32-representable: numpy (f64) and torch (f32) back ends differ
        "frac": ([-0.1, -33.3, -1.7], [47.3, 0.1, 2.9]),
    }
    # put points exactly on the float32 roundings of the fractional bounds
    lo32 = np.asarray(cases["frac"][0], dtype=np.float64).astype(np.float32)
    hi32 = np.asarray(cases["frac"][1], dtype=np.float64).astype(np.float32)
    for k in range(12):
        p[100 + k] = np.float32(1.0)
        p[100 + k, k % 3] = (lo32 if k < 6 else hi32)[k % 3]
    store = {"points": p}
    for cname, (lo, hi) in cases.items():
        store[f"{cname}_min"], store[f"{cname}_max"] = np.array(lo), np.array(hi)
        for backend in ("numpy", "torch"):
            for invert in (False, True):
                cloud = ShimCloud(p.copy())
                utils.crop_pointcloud(cloud, backend=backend, min_bound=lo, max_bound=hi, invert=invert)
                store[f"{cname}_{backend}_{'inv' if invert else 'fwd'}"] = cloud.selected_mask
    np.savez_compressed(os.path.join(out, "crop.npz"), **store)


def gen_dedup(utils, out):
    rng = np.random.default_rng(99)
    n = 3000
    p = rng.uniform(-5, 5, size=(n, 3)).astype(np.float32)
    dst = rng.choice(np.arange(1, n), size=400, replace=False)
    p[dst] = p[(dst * rng.uniform(0, 1, size=400)).astype(int)]
    p[10] = [0.0, 1.0, 2.0]
    p[20] = [-0.0, 1.0, 2.0]          # -0.0 vs +0.0: equal for numpy/torch, distinct bit patterns
    p[30] = [np.nan, 1.0, 2.0]
    p[40] = [np.nan, 1.0, 2.0]        # NaN rows never merge under value equality
    store = {"points": p}
    c = ShimCloud(p.copy())
    utils.remove_duplicates(c, backend="numpy")
    store["numpy_index"] = c.selected_index
    c = ShimCloud(p.copy())
    utils.remove_duplicates(c, backend="torch")
    store["torch_index"] = c.selected_index
    np.savez_compressed(os.path.join(out, "dedup.npz"), **store)


def gen_convert(utils, out):
    rng = np.random.default_rng(7)
    n = 257
    layouts = {
        "velodyne": np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"),
                              ("ring", "<u2"), ("time", "<f4")]),
        "autoware": np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("I", "<u1"), ("R", "<u1"),
                              ("C", "<u2")]),
        "livox": np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("tag", "<u1"),
                           ("line", "<u1"), ("timestamp", "<f8")]),
        "f64xyz": np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("intensity", "<u2")]),
        "rgb": np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgb", "<f4")]),
    }
    store, meta_json = {}, {}
    for name, dt in layouts.items():
        arr = np.zeros(n, dtype=dt)
        for f in dt.names:
            sub = dt.fields[f][0]
            if sub.kind == "f":
                arr[f] = rng.uniform(-100, 100, size=n)
            else:
                arr[f] = rng.integers(0, np.iinfo(sub).max, size=n, endpoint=True)
        if name == "rgb":
            arr["rgb"] = rng.integers(0, 1 << 24, size=n).astype(np.uint32).view(np.float32)
        meta = utils.get_pointcloud_metadata(dt.names)
        meta_full = dict(meta, field_names=dt.names, num_fields=len(dt.names))
        d = utils.convert_pointcloud_to_numpy(arr, meta_full)
        store[f"{name}__bytes"] = np.frombuffer(arr.tobytes(), dtype=np.uint8)
        for k, v in d.items():
            store[f"{name}__{k}"] = v
        meta_json[name] = {"fields": [[f, dt.fields[f][0].str, int(dt.fields[f][1])] for f in dt.names],
                           "itemsize": dt.itemsize, "metadata": meta}
    np.savez_compressed(os.path.join(out, "convert.npz"), **store)
    # field-name mapping on assorted name lists (utils.py:423-472)
    name_lists = [["x", "y", "z"], ["x", "y", "z", "intensity", "ring", "time"],
                  ["x", "y", "z", "I", "R", "C"], ["x", "y", "z", "i", "t", "line", "tag"],
                  ["X", "Y", "Z", "Intensity", "RING"], ["x", "y", "z", "r", "g", "b"],
                  ["x", "y", "z", "rgb", "timestamp", "return_type", "azimuth", "distance"],
                  ["x", "y", "z", "intensity", "I", "time", "t", "timestamp"]]
    meta_json["mappings"] = [{"names": nl, "metadata": utils.get_pointcloud_metadata(nl)} for nl in name_lists]
    # packed output layout (utils.py:140-199)
    packed = []
    for names, dts in ([["x", "y", "z", "intensity", "ring", "time"], [7, 7, 7, 7, 4, 7]],
                       [["x", "y", "z", "I", "R", "C"], [7, 7, 7, 2, 2, 4]],
                       [["x", "y", "z", "t", "d"], [8, 8, 8, 6, 3]]):
        fields, step = utils.numpy_struct_to_pointcloud2(names, dts)
        packed.append({"names": names, "datatypes": dts,
                       "fields": [[f.name, f.offset, f.datatype, f.count] for f in fields], "point_step": step})
    meta_json["packed"] = packed
    # rgb helpers (utils.py:304-356)
    r, g, b = (rng.integers(0, 256, size=64).astype(np.uint8) for _ in range(3))
    merged_f = utils.merge_rgb_fields(r, g, b, return_int=False)
    np.savez_compressed(os.path.join(out, "rgb.npz"), r=r, g=g, b=b, merged_float=merged_f,
                        merged_int=utils.merge_rgb_fields(r, g, b, return_int=True),
                        extracted=utils.extract_rgb_from_pointcloud(merged_f),
                        colors=(np.stack([r, g, b], 1) / 255.0).astype(np.float32),
                        packed=utils.rgb_int_to_float((np.stack([r, g, b], 1) / 255.0).astype(np.float32)),
                        luminance=utils.rgb_to_intensity((np.stack([r, g, b], 1) / 255.0).astype(np.float32)))
    with open(os.path.join(out, "metadata.json"), "w") as f:
        json.dump(meta_json, f, indent=1, default=lambda o: o if not isinstance(o, (np.bool_,)) else bool(o))


def gen_params(out):
    """Parse the node's declare_parameter calls (pp.py:129-199) without importing it."""
    src = open(os.path.join(REF, "autodriver_pointcloud_preprocessor/pointcloud_preprocessor.py")).read()
    tree = ast.parse(src)
    params = []

    def lit(node):
        try:
            return ast.literal_eval(node)
        except Exception:
            if isinstance(node, ast.Call):      # np.eye(4).flatten().tolist() etc.
                return ast.unparse(node)
            return ast.unparse(node)

    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "declare_parameter":
            kw = {k.arg: k.value for k in node.keywords}
            name_node = kw.get("name", node.args[0] if node.args else None)
            value_node = kw.get("value", node.args[1] if len(node.args) > 1 else None)
            name = ast.unparse(name_node)
            # f'{self.parameter_namespace}xyz' -> xyz
            if isinstance(name_node, ast.JoinedStr):
                name = "".join(v.value for v in name_node.values if isinstance(v, ast.Constant))
            ptype = None
            desc = kw.get("descriptor")
            if desc is not None:
                for k in desc.keywords:
                    if k.arg == "type":
                        ptype = ast.unparse(k.value).split(".")[-1]
            params.append({"name": name, "default": lit(value_node), "type": ptype, "line": node.lineno})
    params.sort(key=lambda p: p["line"])
    # processing_times keys written by the node
    keys = sorted({n.slice.value for n in ast.walk(tree)
                   if isinstance(n, ast.Subscript) and ast.unparse(n.value) == "self.processing_times"
                   and isinstance(n.slice, ast.Constant)})
    methods = []
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "PointcloudPreprocessorNode":
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef):
                    methods.append({"name": fn.name, "args": [a.arg for a in fn.args.args]})
    with open(os.path.join(out, "node_contract.json"), "w") as f:
        json.dump({"parameters": params, "processing_times_keys": keys, "methods": methods}, f, indent=1)


def gen_utils_signatures(out):
    src = open(os.path.join(REF, "autodriver_pointcloud_preprocessor/utils.py")).read()
    tree = ast.parse(src)
    sigs = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            a = node.args
            defaults = [None] * (len(a.args) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
            sigs[node.name] = [[arg.arg, d] for arg, d in zip(a.args, defaults)]
    with open(os.path.join(out, "utils_signatures.json"), "w") as f:
        json.dump(sigs, f, indent=1)


def main():
    utils = load_reference_utils()
    gen_crop(utils, HERE)
    gen_dedup(utils, HERE)
    gen_convert(utils, HERE)
    gen_params(HERE)
    gen_utils_signatures(HERE)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
