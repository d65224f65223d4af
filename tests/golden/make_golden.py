"""Regenerates tests/golden/ from the read-only reference checkout (run in the build
container only; /root/reference does not exist on the GPU box).

Copies the small example decks + golden outputs the reference ships for the p121/p123
path and derives compact fixtures from the large ones:
  xx3-tiny.*            copied (125 hex20 deck, .res 79 iterations, .dis displacements)
  p121_demo.{mg,dat,lds,res}  copied; the 4 MB .d / .bnd are replaced by SHA-256 digests of
                        the parsed arrays (the in-memory generator must reproduce them) and
                        the EnSight displacement golden by a compressed .npz of its values
  p121_book.{mg,res}, p123_book.{mg,res}, p123_small.mg   copied
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference/parafem/examples"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from parafem_b200 import host
    for f in ("bnd", "d", "dat", "dis", "lds", "res"):
        shutil.copy(f"{REF}/dev/xx3/demo/xx3-tiny.{f}", f"{HERE}/xx3-tiny.{f}")
    for f in ("mg", "dat", "lds", "res"):
        shutil.copy(f"{REF}/5th_ed/p121/demo/p121_demo.{f}", f"{HERE}/p121_demo.{f}")
    shutil.copy(f"{REF}/5th_ed/p121/book/p121.mg", f"{HERE}/p121_book.mg")
    shutil.copy(f"{REF}/5th_ed/p121/book/p121.res", f"{HERE}/p121_book.res")
    shutil.copy(f"{REF}/5th_ed/p123/book/p123.mg", f"{HERE}/p123_book.mg")
    shutil.copy(f"{REF}/5th_ed/p123/book/p123.res", f"{HERE}/p123_book.res")
    shutil.copy(f"{REF}/5th_ed/p123/mg/p123_small.mg", f"{HERE}/p123_small.mg")
    shutil.copy(f"{REF}/5th_ed/p121/mg/p121_tiny.mg", f"{HERE}/p121_tiny.mg")
    for f in os.listdir(HERE):
        os.chmod(os.path.join(HERE, f), 0o644)
    p = host.read_deck_p121(f"{REF}/5th_ed/p121/demo/p121_demo")
    digests = dict(g_num_sg=sha(p.g_num_pp), g_coord_pp=sha(p.g_coord_pp), rest=sha(p.rest), g_g=sha(p.g_g_pp),
                   r=sha(p.r_pp), nn=int(p.nn), nr=int(p.nr), neq=int(p.neq), nels=int(p.nels))
    json.dump(digests, open(f"{HERE}/p121_demo_digests.json", "w"), indent=1)
    disp = np.loadtxt(f"{REF}/5th_ed/p121/demo/p121_demo.ensi.DISPL-000001", skiprows=4)
    np.savez_compressed(f"{HERE}/p121_demo_displ.npz", displ=disp.reshape(3, p.nn).T.astype(np.float32))
    with open(f"{REF}/5th_ed/p121/demo/p121_demo.ensi.DISPL-000001") as f, open(f"{HERE}/p121_demo_ensi_head.txt", "w") as g:
        g.writelines([next(f) for _ in range(204)])   # header + the first 200 x-displacements, verbatim
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
