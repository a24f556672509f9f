#!/usr/bin/env python
"""Generate tests/golden/bam/*.gz: outputs of the UNMODIFIED reference binary (oracle/_ref/bsmap) fed with BAM read
files (reads.cpp:120-143).  The BAM inputs are rebuilt deterministically by tests/bam_cases.py from the parity cases.
    make -C oracle ref && python tests/golden/make_bam_golden.py"""
import gzip
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import bam_cases as BC   # noqa: E402
import oracle_lib as O   # noqa: E402

OUT = os.path.join(HERE, "bam")


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for name in BC.NAMES:
        with tempfile.TemporaryDirectory() as td:
            argv, out = BC.build(name, td)
            stdout = O.run_reference(argv + ["-p", "1"], cwd=td)
            txt = open(out, "rb").read()
        with gzip.GzipFile(os.path.join(OUT, name + ".sam.gz"), "wb", mtime=0) as f:
            f.write(txt)
        manifest[name] = dict(lines=txt.count(b"\n"), summary=[l for l in stdout.splitlines() if "aligned" in l or l.startswith(("pairs", "single"))])
        print(name, manifest[name])
    json.dump(manifest, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
