"""Memory safety of the htslib-free readers on damaged input: the text VCF parser, the BCF2.2 decoder (below the BGZF layer)
and the front half of the `call -m` driver (b200_vc_open / b200_vc_next) are fed mutated copies of the reference's test files
under AddressSanitizer + UBSan (tests/fuzz/*.c, host code only: the device library is stubbed out).  Damaged input must come
back as an error, never as a crash; the first run of this harness found a negative typed-vector length in
b200_bcf_decode_rec reaching malloc()."""
import os
import shutil
import subprocess

import pytest

from tests import vcf_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "bcftools_b200", "csrc", "host")
FUZZ = os.path.join(ROOT, "tests", "fuzz")
SAN = ["-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I", os.path.join(ROOT, "include")]


def _build(tmp, name, sources):
    exe = os.path.join(tmp, name)
    r = subprocess.run(["gcc"] + SAN + ["-o", exe] + sources + ["-lz", "-lm", "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0 and "asan" in (r.stderr or "").lower():
        pytest.skip("no AddressSanitizer runtime in this image")
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def _fixture(tmp, name, nrec):
    files = vcf_cases.bundle()["files"]
    lines = files[name].split("\n")
    text = "\n".join([l for l in lines if l.startswith("#")] + [l for l in lines if l and not l.startswith("#")][:nrec]) + "\n"
    path = os.path.join(tmp, name)
    open(path, "w").write(text)
    return path


def _run(exe, args, cwd):
    r = subprocess.run([exe] + args, capture_output=True, text=True, cwd=cwd, timeout=600, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert r.returncode == 0 and "Sanitizer" not in r.stderr and "runtime error" not in r.stderr, (r.stdout[-500:], r.stderr[-3000:])
    return r.stdout


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_vcf_and_bcf_readers_survive_damaged_input(tmp_path):
    tmp = str(tmp_path)
    exe = _build(tmp, "fuzz_readers", [os.path.join(FUZZ, "fuzz_vcf_bcf_readers.c"), os.path.join(HOST, "b200_vcf.c"), os.path.join(HOST, "b200_bcfio.c")])
    for name, seed in (("mpileup.cals.1.vcf", 7), ("call-G.vcf", 8)):
        out = _run(exe, [_fixture(tmp, name, 12), "250", str(seed)], tmp)
        assert "bcf ok" in out and "rejected" in out


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_call_driver_front_half_survives_damaged_input(tmp_path):
    tmp = str(tmp_path)
    src = [os.path.join(FUZZ, "fuzz_call_driver.c"), os.path.join(FUZZ, "device_stubs.c")] + \
          [os.path.join(HOST, f) for f in ("b200_vcf.c", "b200_bcfio.c", "b200_vcfcall.c", "b200_driver.c", "b200_pv4.c", "b200_bcf.c", "b200_call.c")]
    exe = _build(tmp, "fuzz_driver", src)
    files = vcf_cases.bundle()["files"]
    for aux in ("mpileup.ploidy", "mpileup.2.samples", "mpileup.cals.2.tab", "call.af-fixation.txt"):
        open(os.path.join(tmp, aux), "w").write(files[aux])
    for name, nrec, seed, args in (("mpileup.vcf", 20, 1, ["-mv"]),
                                   ("mpileup.X.vcf", 20, 3, ["-mv", "--ploidy-file", "mpileup.ploidy", "-S", "mpileup.2.samples"]),
                                   ("call.af-fixation.vcf", 10, 7, ["-m", "-G", "call.af-fixation.txt", "-a", "GP,GQ"]),
                                   ("mpileup.cals.2.vcf", 20, 8, ["-mA", "-C", "alleles", "-T", "mpileup.cals.2.tab", "-i"])):
        out = _run(exe, [_fixture(tmp, name, nrec), "150", str(seed)] + args, tmp)
        assert out.startswith("opened")
