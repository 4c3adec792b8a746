"""Writes the INPUTS of the golden cases (tests/golden/*.npz) that the reference itself can run -- PCD at any batch size and
online CD (batch size 1; mini-batch CD is this repository's declared extension) -- in a format Julia reads without any
package, for tests/golden/make_golden_from_reference.jl:

    tests/golden/reference_inputs/<case>/manifest.txt     one line per entry:  name kind dims... | scalar value
    tests/golden/reference_inputs/<case>/<name>.f64       raw little-endian Float64, COLUMN-major with the listed dims
                                                          (reshape(data, dims...) in Julia indexes like the numpy array)

The random streams are exported in the reference's SERIAL consumption order: `su_<t>` are the
uniforms `rand()` returns during mini-batch step t, `sz_<t>` the standard normals behind `rand(Normal(mu, 1.0))`.

    python tests/golden/export_reference_inputs.py          # rewrites tests/golden/reference_inputs (deterministic)
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.helpers import DenseStreams  # noqa: E402


def reference_runnable(d) -> bool:
    return str(d["algo"]) == "pcd" or int(d["B"]) == 1


def write_case(path_npz: str, out_root: str) -> str | None:
    g = np.load(path_npz)
    d = {k: g[k] for k in g.files}
    if not reference_runnable(d):
        return None
    name = os.path.basename(path_npz)[:-4]
    out = os.path.join(out_root, name)
    os.makedirs(out, exist_ok=True)
    lines = []

    def scalar(key, value):
        lines.append(f"{key} scalar {value}")

    def array(key, a):
        a = np.asarray(a, dtype=np.float64)
        np.asfortranarray(a).ravel(order="F").astype("<f8").tofile(os.path.join(out, key + ".f64"))
        lines.append(f"{key} array " + " ".join(str(int(s)) for s in a.shape))

    V, H, L, B, k, steps = (int(d[x]) for x in ("V", "H", "L", "B", "k", "steps"))
    gaussian, algo = bool(d["gaussian"]), str(d["algo"])
    for key in ("V", "H", "L", "B", "k", "steps"):
        scalar(key, int(d[key]))
    scalar("gaussian", int(gaussian)); scalar("algo", algo); scalar("lr", repr(float(d["lr"]))); scalar("llr", repr(float(d["llr"])))
    array("W0", d["W0"]); array("a0", d["a0"]); array("b0", d["b0"]); array("X", d["X"])
    if L:
        array("U0", d["U0"]); array("c0", d["c0"]); array("Y", d["Y"])
    if algo == "pcd":
        array("cv0", d["cv0"]); array("ch0", d["ch0"])
        if L:
            array("cy0", d["cy0"])
    for t in range(steps):
        ds = DenseStreams(np.random.default_rng(0), k, B, V, H, L, gaussian)
        ds.U_h, ds.U_v, ds.U_y, ds.Z_v = d[f"Uh_{t}"], d.get(f"Uv_{t}"), d.get(f"Uy_{t}"), d.get(f"Zv_{t}")
        su, sz = ds.serial(algo)
        array(f"su_{t}", su)
        if sz is not None:
            array(f"sz_{t}", sz)
    with open(os.path.join(out, "manifest.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return name


def read_case(path_dir: str) -> dict:
    """Reader for the same container (inputs and the reference's outputs), used by tests/test_golden.py."""
    out = {}
    with open(os.path.join(path_dir, "manifest.txt")) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[1] == "scalar":
                out[p[0]] = p[2]
            else:
                dims = tuple(int(x) for x in p[2:])
                a = np.fromfile(os.path.join(path_dir, p[0] + ".f64"), dtype="<f8")
                out[p[0]] = a.reshape(dims, order="F") if dims else a.reshape(())
    return out


def main(out_root: str | None = None):
    out_root = out_root or os.path.join(HERE, "reference_inputs")
    done = []
    for p in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        n = write_case(p, out_root)
        if n:
            done.append(n)
    return done


if __name__ == "__main__":
    for n in main(sys.argv[1] if len(sys.argv) > 1 else None):
        print("wrote", n)
