#!/usr/bin/env python
"""Generate tests/golden/*.gz: outputs of the UNMODIFIED reference binary (oracle/_ref/bsmap,
built from /root/reference by `make -C oracle ref`) for every parity case in tests/cases.py.

Run in the build container (the reference sources do not exist on the GPU box):
    make -C oracle ref && python tests/golden/make_golden.py
Inputs are regenerated from seeds; MANIFEST.json records their sha256 so a drifting generator is
caught by tests/test_oracle_vs_golden.py.
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases as CS      # noqa: E402
import runners as R     # noqa: E402


def main():
    manifest = {}
    sel = sys.argv[1:]
    mpath = os.path.join(HERE, "MANIFEST.json")
    if sel and os.path.exists(mpath):
        manifest = json.load(open(mpath))
    for c in CS.CASES:
        if sel and c.name not in sel:
            continue
        main_txt, un_txt, stdout = R.reference_run(c)
        m, u = R.golden_paths(c)
        with gzip.GzipFile(m, "wb", mtime=0) as f:
            f.write(main_txt)
        if un_txt:
            with gzip.GzipFile(u, "wb", mtime=0) as f:
                f.write(un_txt)
        summary = [l for l in stdout.splitlines() if "aligned" in l or l.startswith(("pairs", "single"))]
        manifest[c.name] = dict(inputs_sha256=CS.input_digest(c), cli=" ".join(c.cli("a", "b", "ref.fa", "out." + c.out_ext,
                                "out_unpair.bsp" if (c.paired and c.out_ext != "sam") else None)),
                                lines=main_txt.count(b"\n"), unpair_lines=un_txt.count(b"\n"), summary=summary)
        print(c.name, manifest[c.name]["lines"], summary)
    json.dump(manifest, open(mpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
