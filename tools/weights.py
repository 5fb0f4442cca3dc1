"""Reader / writer for the reference's weight files (SaveLoad_conv, netlib.cpp:220-272; aefft_saveload_conv writes the
same bytes): ./weights/C_weights_{L}{_in|_out}_D={dD}_M={dM}_Lk={Lk}_Ll={Ll}_S={scale}.conv holds raw little-endian
float32, the kernel tensor c[dM][dD][Nk][Nl] (Nk = 2(Lk+1)+1) followed by the bias b[dM]; every dimension lives in the
file NAME only.  For analysis outside the engine:

    python tools/weights.py ./weights            # one line per file: shape, |c|, min/max, bias range
    from weights import read_conv; c, b, meta = read_conv(path)
"""
import os
import re
import sys

import numpy as np

NAME = re.compile(r"C_weights_(?P<L>\d+)_(?P<io>in|out)_D=(?P<dD>\d+)_M=(?P<dM>\d+)_Lk=(?P<Lk>-?\d+)_Ll=(?P<Ll>-?\d+)_S=(?P<S>-?\d+)\.conv$")


def parse_name(path):
    """dict(L, io, dD, dM, Lk, Ll, scale, Nk, Nl) from a weight-file name; ValueError if it is not one."""
    m = NAME.search(os.path.basename(str(path)))
    if not m:
        raise ValueError(f"not a SaveLoad_conv file name: {path}")
    d = {k: int(v) for k, v in m.groupdict().items() if k != "io"}
    d["io"] = 0 if m.group("io") == "in" else 1
    d["scale"] = d.pop("S")
    d["Nk"], d["Nl"] = 2 * (d["Lk"] + 1) + 1, 2 * (d["Ll"] + 1) + 1
    return d


def file_name(L, io, dD, dM, Nk, Nl, scale):
    return f"C_weights_{L}{'_in' if io == 0 else '_out'}_D={dD}_M={dM}_Lk={(Nk - 1) // 2 - 1}_Ll={(Nl - 1) // 2 - 1}_S={scale}.conv"


def read_conv(path):
    """(c[dM][dD][Nk][Nl], b[dM], meta) -- raises ValueError on a size that does not match the name (the reference would
    silently read short)."""
    meta = parse_name(path)
    n_c = meta["dM"] * meta["dD"] * meta["Nk"] * meta["Nl"]
    raw = np.fromfile(str(path), dtype="<f4")
    if raw.size != n_c + meta["dM"]:
        raise ValueError(f"{path}: {raw.size} floats, expected {n_c} + {meta['dM']}")
    return raw[:n_c].reshape(meta["dM"], meta["dD"], meta["Nk"], meta["Nl"]).copy(), raw[n_c:].copy(), meta


def write_conv(directory, c, b, scale, L, io):
    c = np.ascontiguousarray(c, dtype="<f4")
    b = np.ascontiguousarray(b, dtype="<f4")
    dM, dD, Nk, Nl = c.shape
    assert b.shape == (dM,)
    path = os.path.join(str(directory), file_name(L, io, dD, dM, Nk, Nl, scale))
    with open(path, "wb") as fh:
        fh.write(c.tobytes())
        fh.write(b.tobytes())
    return path


MOM = re.compile(r"C_momentum_(?P<L>\d+)_D=(?P<dD>\d+)_M=(?P<dM>\d+)_Lk=(?P<Lk>-?\d+)_Ll=(?P<Ll>-?\d+)\.mom$")


def read_momentum(path):
    """The engine's momentum sidecar (aefft_net_saveload_momentum; the reference does not save this state):
    dict(dc, db, df, dp, ddc, ddb, ddf, ddp) + meta."""
    m = MOM.search(os.path.basename(str(path)))
    if not m:
        raise ValueError(f"not a momentum sidecar name: {path}")
    meta = {k: int(v) for k, v in m.groupdict().items()}
    Nk, Nl = 2 * (meta["Lk"] + 1) + 1, 2 * (meta["Ll"] + 1) + 1
    dM, dD = meta["dM"], meta["dD"]
    nC = dM * dD * Nk * Nl
    raw = np.fromfile(str(path), dtype="<f4")
    if raw.size != 4 * nC + 2 * (dM + dD):
        raise ValueError(f"{path}: {raw.size} floats, expected {4 * nC + 2 * (dM + dD)}")
    out, off = {}, 0
    for name, n, shape in (("dc", nC, (dM, dD, Nk, Nl)), ("db", dM, (dM,)), ("df", nC, (dD, dM, Nk, Nl)), ("dp", dD, (dD,)),
                           ("ddc", nC, (dM, dD, Nk, Nl)), ("ddb", dM, (dM,)), ("ddf", nC, (dD, dM, Nk, Nl)), ("ddp", dD, (dD,))):
        out[name] = raw[off:off + n].reshape(shape).copy()
        off += n
    return out, meta


def main(argv):
    directory = argv[1] if len(argv) > 1 else "./weights"
    names = sorted(n for n in os.listdir(directory) if NAME.search(n))
    if not names:
        print(f"no weight files under {directory}")
        return 1
    for n in names:
        c, b, m = read_conv(os.path.join(directory, n))
        print(f"pair {m['L']} {'encoder' if m['io'] == 0 else 'decoder'}: {m['dD']:>3} -> {m['dM']:<3} {m['Nk']}x{m['Nl']} taps, pool {m['scale']:>2}"
              f" | |c| {np.linalg.norm(c):.4g}  c in [{c.min():.4g}, {c.max():.4g}]  b in [{b.min():.4g}, {b.max():.4g}]")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
