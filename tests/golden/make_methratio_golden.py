#!/usr/bin/env python
"""Generate tests/golden/methratio/*.gz: outputs of the reference's methratio.py on the golden alignment files.

methratio.py is Python 2 and shells out to `samtools view -X[S]` for SAM input.  Neither exists in this image,
so the script is run from /root/reference through three MECHANICAL adaptations (nothing of its logic is touched):
  * `print >> sys.stderr, x` / `print x`  ->  print(...)      * `xrange` -> `range`
  * `os.popen('<samtools> view -XS file')` -> a reader that yields the SAM body with FLAG rendered the way
    samtools 0.1.x `-X` renders it (one letter per set bit: p P u U r R 1 2 s f d for 0x1 ... 0x400).
BAM input (BAM_RUNS) goes through the REAL vendored samtools built by `make -C oracle ref`: the golden SAM is turned
into a sorted BAM by `samtools view -bS | samtools sort` (what the reference's sam2bam.sh does) and the script reads it
through its own `os.popen('samtools view -X file.bam')`, untouched.
Run in the build container:  make -C oracle ref && python tests/golden/make_methratio_golden.py
The adapted source is never written to the repo; only its outputs are (as fixtures), with this script.
"""
import gzip
import io
import json
import os
import re
import sys
import tempfile
from contextlib import redirect_stderr, redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases as CS      # noqa: E402
import runners as R     # noqa: E402

REF = "/root/reference/methratio.py"
OUT = os.path.join(HERE, "methratio")
FLAG_LETTERS = "pPuUrR12sfd"

# (case, methratio options) -- every option the script has except -s (samtools path)
RUNS = [
    ("se_cfg2_r0_uR", []), ("se_cfg2_r0_uR", ["-z", "-m", "2"]), ("se_cfg2_bsp", []), ("se_cfg2_bsp", ["-u", "-t", "0"]),
    ("se_cfg1", ["-u"]), ("se_cfg1", ["-g", "-z"]), ("se_n1", []), ("se_n1", ["-t", "5", "-g"]), ("se_mixed_A", ["-z"]),
    ("pe_sam", []), ("pe_sam", ["-p", "-u"]), ("pe_sam_v5_R", ["-t", "3", "-z"]), ("pe_readthrough", []), ("pe_readthrough", ["-p", "-g"]),
    ("pe_bsp_r0", []), ("pe_bsp_r0", ["-p"]), ("pe_n1", ["-c", "chr2,chr1"]), ("rrbs_se_A", ["-z"]), ("rrbs_pe", ["-t", "0"]),
    ("se_cfg5", ["-m", "3"]),
    # -r: the first alignment (file order) of each (chromosome, fragment end, direction) counts
    ("se_cfg2_r0_uR", ["-r"]), ("se_cfg2_bsp", ["-r", "-u"]), ("se_cfg5", ["-r", "-z"]), ("se_n1", ["-r", "-t", "0"]), ("pe_sam", ["-r"]),
    ("pe_sam_v5_R", ["-r", "-p"]), ("pe_readthrough", ["-r", "-g"]), ("pe_bsp_r0", ["-r"]), ("rrbs_se_A", ["-r"]), ("rrbs_pe", ["-r", "-u"]),
]


# (case, methratio options) on the case's golden SAM converted to a sorted BAM; -r then follows the sorted order
BAM_RUNS = [("se_cfg2_r0_uR", []), ("pe_sam", ["-r"]), ("rrbs_se_A", ["-r", "-u"]), ("pe_readthrough", ["-p", "-g"]), ("rrbs_pe", ["-r", "-t", "0"])]
SAMTOOLS_DIR = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref")


def samtools_view_X(path):
    """the SAM body (no header) with the FLAG column as samtools 0.1.x -X prints it"""
    out = io.StringIO()
    for line in open(path):
        if line.startswith("@"):
            continue
        col = line.split("\t")
        f = int(col[1])
        col[1] = "".join(ch for k, ch in enumerate(FLAG_LETTERS) if f & (1 << k))
        out.write("\t".join(col))
    out.seek(0)
    return out


def adapted_source():
    src = open(REF).read()
    src = src.replace("xrange", "range")
    src = re.sub(r"print >> sys\.stderr, (.*)$", r"print(\1, file=sys.stderr)", src, flags=re.M)
    src = re.sub(r"^print (.*)$", r"print(\1)", src, flags=re.M)
    src = src.replace("os.popen('%ssamtools view -XS %s' % (options.sam_path, infile))", "samtools_view_X(infile)")
    # the .BAM branch keeps its os.popen('%ssamtools view -X %s'): BAM_RUNS pass -s <oracle/_ref>
    return src


def run_reference(argv):
    """exec the adapted script with argv; returns its stdout"""
    code = compile(adapted_source(), REF, "exec")
    old = sys.argv
    sys.argv = ["methratio.py"] + argv
    so, se = io.StringIO(), io.StringIO()
    try:
        with redirect_stdout(so), redirect_stderr(se):
            exec(code, {"__name__": "__main__", "samtools_view_X": samtools_view_X})
    finally:
        sys.argv = old
    return so.getvalue()


def tag(opts):
    return "default" if not opts else "_".join(o.lstrip("-").replace(",", "+") for o in opts)


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for name, opts in RUNS:
        case = CS.BY_NAME[name]
        with tempfile.TemporaryDirectory() as td:
            fa, _, _ = CS.write_inputs(case, td)
            main_txt, un_txt = R.golden_load(case)
            aln = os.path.join(td, "aln." + case.out_ext)
            open(aln, "wb").write(main_txt)
            files = [aln]
            if un_txt:
                files.append(os.path.join(td, "aln_unpair.bsp"))
                open(files[-1], "wb").write(un_txt)
            out = os.path.join(td, "meth.txt")
            stdout = run_reference(["-o", out, "-d", fa, "-q"] + opts + files)
            txt = open(out, "rb").read()
        key = f"{name}.{tag(opts)}"
        with gzip.GzipFile(os.path.join(OUT, key + ".txt.gz"), "wb", mtime=0) as f:
            f.write(txt)
        manifest[key] = dict(case=name, opts=opts, lines=txt.count(b"\n"), stdout=stdout.strip())
        print(key, manifest[key]["lines"], stdout.strip())
    import subprocess
    for name, opts in BAM_RUNS:
        case = CS.BY_NAME[name]
        with tempfile.TemporaryDirectory() as td:
            fa, _, _ = CS.write_inputs(case, td)
            main_txt, _ = R.golden_load(case)
            sam = os.path.join(td, "aln.sam")
            open(sam, "wb").write(main_txt)
            st = os.path.join(SAMTOOLS_DIR, "samtools")
            subprocess.run(f"{st} view -bS {sam} > {td}/tmp.bam 2>/dev/null && {st} sort {td}/tmp.bam {td}/aln", shell=True, check=True)
            out = os.path.join(td, "meth.txt")
            stdout = run_reference(["-o", out, "-d", fa, "-q", "-s", SAMTOOLS_DIR] + opts + [os.path.join(td, "aln.bam")])
            txt = open(out, "rb").read()
        key = f"{name}.bam.{tag(opts)}"
        with gzip.GzipFile(os.path.join(OUT, key + ".txt.gz"), "wb", mtime=0) as f:
            f.write(txt)
        manifest[key] = dict(case=name, opts=opts, lines=txt.count(b"\n"), stdout=stdout.strip(), bam=True)
        print(key, manifest[key]["lines"], stdout.strip())
    json.dump(manifest, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
