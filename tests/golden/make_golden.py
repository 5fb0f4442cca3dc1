"""Generates tests/golden/cpu_golden.npz from the UNMODIFIED reference (oracle/_ref/libref.so, built by
oracle/Makefile from /root/reference/source/netlib.cpp).  CPU functions only, so it runs in the build container:

    make -C oracle && python tests/golden/make_golden.py

Inputs and reference outputs are stored together; tests/test_oracle_cpu.py replays them through oracle/oracle_np.py.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_lib as R  # noqa: E402


def main():
    g = {}
    rng = np.random.default_rng(7)
    # Init_conv under srand(1234): the bench's seeded weights
    R.srand(1234)
    c, b = R.init_conv(4, 3, 5, 5, 3.0)
    f, p = R.init_conv(3, 4, 5, 5, 3.0)
    g.update(init_c=c, init_b=b, init_f=f, init_p=p)
    # Pool down / up
    x = (rng.random((3, 12, 10)) * 300 - 40).astype(np.float32)
    g["pool_x"] = x
    g["pool_down2"] = R.pool(x, 2, (6, 5))
    g["pool_down1"] = R.pool(x, 1, (12, 10))
    g["pool_up2"] = R.pool(g["pool_down2"], -2, (12, 10))
    # Conv (CPU convention) 5x5 and 3x3
    for tag, (dM, dD, Nk, Nl, Nx, Ny) in {"a": (4, 3, 5, 5, 14, 11), "b": (2, 2, 3, 3, 9, 9), "c": (3, 1, 7, 5, 16, 12)}.items():
        xin = (rng.random((dD, Nx, Ny)) * 255).astype(np.float32)
        cc = (rng.random((dM, dD, Nk, Nl)) * 2 - 1).astype(np.float32)
        bb = (rng.random(dM) * 2 - 1).astype(np.float32)
        g[f"conv_{tag}_x"], g[f"conv_{tag}_c"], g[f"conv_{tag}_b"] = xin, cc, bb
        g[f"conv_{tag}_out"] = R.conv_cpu(xin, cc, bb)
    # backprop (CPU): sequential-f semantics; config-1-like (D=1,M=8,5x5) small frame + a D=3 case
    for tag, (dM, dD, Nk, Nl, Nx, Ny, delta) in {"c1": (8, 1, 5, 5, 20, 16, 0.2), "d3": (4, 3, 5, 5, 14, 12, 1.0),
                                                  "k3": (3, 2, 3, 3, 10, 10, 0.5)}.items():
        inp = (rng.random((dD, Nx, Ny)) * 255).astype(np.float32)
        cc = ((rng.random((dM, dD, Nk, Nl)) * 2 - 1) * 0.2).astype(np.float32)
        ff = ((rng.random((dD, dM, Nk, Nl)) * 2 - 1) * 0.2).astype(np.float32)
        bb = (rng.random(dM) * 2 - 1).astype(np.float32)
        pp = (rng.random(dD) * 2 - 1).astype(np.float32)
        hin = R.conv_cpu(inp, cc, bb)
        out = R.conv_cpu(hin, ff, pp)
        res = R.backprop_cpu(inp, out, hin, cc, bb, ff, pp, delta)
        for k, v in dict(inp=inp, hin=hin, out=out, c=cc, b=bb, f=ff, p=pp, delta=np.float32(delta)).items():
            g[f"bp_{tag}_{k}"] = v
        for k, v in res.items():
            g[f"bp_{tag}_new_{k}"] = v
    # Portion
    a, h, o = R.portion(g["bp_d3_inp"], g["bp_d3_hin"], g["bp_d3_out"], 2)
    g.update(portion_in=a, portion_hin=h, portion_out=o)
    # kernel_pad
    g["kpad_c"] = g["conv_a_c"]
    g["kpad_out"] = R.kernel_pad(g["conv_a_c"], 16, 8)
    # SaveLoad_conv byte format: written by the reference into ./weights/
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "weights"))
        os.chdir(td)
        try:
            R.saveload_conv(g["conv_a_c"], g["conv_a_b"], 2, 0, 0, 1)
            names = os.listdir("weights")
            assert len(names) == 1
            g["save_name"] = np.array(names[0])
            g["save_bytes"] = np.frombuffer(open(os.path.join("weights", names[0]), "rb").read(), np.uint8)
        finally:
            os.chdir(cwd)
    out = os.path.join(ROOT, "tests", "golden", "cpu_golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
