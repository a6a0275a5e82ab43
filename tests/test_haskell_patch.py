"""The reference-side edit (haskell/patch/*.patch + the FFI module) is committed as files, not prose.  GHC is not in the
image, so nothing here compiles Haskell: the tests check that the patches APPLY to the reference tree, that patch 0001
(the drop-in) touches only the lines SURVEY.md Appendix E lists, and that every `FFT.*` name the patched modules use is
exported by B200FFT.hs with every C symbol it imports declared in include/b200fft.h.  Skipped where /root/reference is
absent (the GPU box)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
P1 = os.path.join(ROOT, "haskell/patch/0001-ptx-backend-calls-libb200fft.patch")
P2 = os.path.join(ROOT, "haskell/patch/0002-fused-inverse-and-any-rank.patch")
FFI = os.path.join(ROOT, "haskell/Data/Array/Accelerate/Math/FFT/LLVM/PTX/B200FFT.hs")

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")) or shutil.which("patch") is None,
                               reason="reference tree or patch(1) not available")

# the reference lines the drop-in may touch (inclusive ranges, reference numbering)
ALLOWED_0001 = {
    "accelerate-fft.cabal": [(82, 98)],
    "src/Data/Array/Accelerate/Math/FFT/LLVM/PTX.hs": [(44, 45), (77, 78), (112, 124)],
    "src/Data/Array/Accelerate/Math/FFT/LLVM/PTX/Plans.hs": [(29, 33), (41, 42), (66, 66), (72, 72)],
}


def _tree(tmp_path):
    dst = tmp_path / "ref"
    dst.mkdir()
    shutil.copytree(os.path.join(REF, "src"), dst / "src")
    shutil.copy(os.path.join(REF, "accelerate-fft.cabal"), dst / "accelerate-fft.cabal")
    return dst


def _apply(tree, patch):
    r = subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-i", patch], cwd=tree, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "fuzz" not in r.stdout and "offset" not in r.stdout, r.stdout    # hunks apply at their stated lines


def _touched(patch):
    """{file: [reference line numbers removed or followed by an insertion]} from a unified diff"""
    out, cur, old = {}, None, 0
    for line in open(patch):
        if line.startswith("--- a/"):
            cur = line[6:].strip()
            out[cur] = []
        elif line.startswith("+++ ") or cur is None:
            continue
        elif line.startswith("@@"):
            old = int(re.match(r"@@ -(\d+)", line).group(1))
        elif line.startswith("-"):
            out[cur].append(old)
            old += 1
        elif line.startswith("+"):
            out[cur].append(old)        # an insertion sits before reference line `old`
        elif line.startswith(" "):
            old += 1
    return out


@needs_ref
def test_patches_apply_to_the_reference_tree(tmp_path):
    tree = _tree(tmp_path)
    _apply(tree, P1)
    ptx = (tree / "src/Data/Array/Accelerate/Math/FFT/LLVM/PTX.hs").read_text()
    assert "Foreign.CUDA.FFT" not in ptx and "setStream" not in ptx
    library = (tree / "accelerate-fft.cabal").read_text().split("test-suite")[0]
    assert "cufft" not in library.split("build-depends")[-1] and "extra-libraries:    b200fft" in library
    _apply(tree, P2)
    top = (tree / "src/Data/Array/Accelerate/Math/FFT.hs").read_text()
    assert top.count("foreignAcc (PTX.") == 4 and "rank P.<= 3" not in top
    # nothing outside the hot path was touched
    for rel in ("Mode.hs", "Type.hs", "Adhoc.hs", "LLVM/Native.hs", "LLVM/PTX/Base.hs"):
        a = open(os.path.join(REF, "src/Data/Array/Accelerate/Math/FFT", rel)).read()
        assert (tree / "src/Data/Array/Accelerate/Math/FFT" / rel).read_text() == a, rel


@needs_ref
def test_drop_in_patch_touches_only_the_listed_lines():
    touched = _touched(P1)
    assert set(touched) == set(ALLOWED_0001), sorted(touched)
    for f, lines in touched.items():
        for ln in lines:
            assert any(lo <= ln <= hi + 1 for lo, hi in ALLOWED_0001[f]), (f, ln)


def test_ffi_module_covers_the_patched_call_sites_and_the_header():
    exports = re.search(r"module [\w.]+ \((.*?)\) where", open(FFI).read(), re.S).group(1)
    exported = set(re.findall(r"\b([A-Za-z][\w']*)", re.sub(r"\(\.\.\)", "", exports)))
    used = set()
    for p in (P1, P2):
        for line in open(p):
            if line.startswith("+") and not line.startswith("+++"):
                used |= set(re.findall(r"(?<![\w.])FFT\.([A-Za-z][\w']*)", line)) - {"LLVM", "hs"}
    assert used and used <= exported, sorted(used - exported)
    header = open(os.path.join(ROOT, "include/b200fft.h")).read()
    for sym in re.findall(r'foreign import ccall \w+\s+"(\w+)"', open(FFI).read()):
        assert re.search(r"\b%s\s*\(" % sym, header), sym
