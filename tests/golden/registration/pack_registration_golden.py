"""registration_golden.txt (output of make_registration_golden, the REAL reference) -> tests/golden/registration_golden.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def parse(path):
    tok = open(path).read().split()
    pos = 0

    def take(n=1):
        nonlocal pos
        v = tok[pos:pos + n]
        pos += n
        return v

    assert take()[0] == "cases"
    out = {}
    names = []
    for _ in range(int(take()[0])):
        assert take()[0] == "case"
        name = take()[0]
        names.append(name)
        assert take()[0] == "result"
        out[f"{name}/result"] = np.array(take(7), dtype=np.float64)
        assert take()[0] == "termination"
        out[f"{name}/termination"] = np.int32(take()[0])
        assert take()[0] == "iterations"
        n_it = int(take()[0])
        est, upd = np.zeros((n_it, 7)), np.zeros((n_it, 7))
        for i in range(n_it):
            assert take()[0] == "est"
            est[i] = np.array(take(7), dtype=np.float64)
            assert take()[0] == "update"
            upd[i] = np.array(take(7), dtype=np.float64)
            for kind in ("edge_assoc", "plane_assoc"):
                assert take()[0] == kind
                n = int(take()[0])
                out[f"{name}/{kind}/{i}"] = np.array(take(2 * n), dtype=np.uint32).reshape(n, 2)
        out[f"{name}/iter_est"], out[f"{name}/iter_update"] = est, upd
    out["names"] = np.array(names)
    return out


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "registration_golden.txt")
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(HERE), "registration_golden.npz")
    np.savez_compressed(dst, **parse(src))
    print("wrote", dst)
