"""The reference-facing boundary (dropin/*.cpp: definitions of the reference's own ORBextractor / ORBMatcher / Optimizer / System methods
on the C ABI) must compile against the reference's UNMODIFIED headers. This container has no OpenCV C++, so `cv::` comes from the shim in
oracle/cvshim, whose _InputArray / _OutputArray are proxy objects with getMat() / create() / release() like OpenCV's (the facade of round 1
read `.rows` / `.data` on them directly and could not have compiled against real OpenCV). Needs the reference tree: build container only;
the link-and-run half of this check is tests/test_gpu_dropin.py on the GPU box."""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")


@pytest.mark.parametrize("src", sorted(glob.glob(os.path.join(ROOT, "dropin", "*.cpp"))))
def test_dropin_compiles_against_reference_headers(src):
    gxx = shutil.which("g++")
    if gxx is None or not os.path.isdir(REF):
        pytest.skip("needs g++ and the reference tree")
    cmd = [gxx, "-std=c++11", "-fsyntax-only", "-Wall", "-Wno-unused-function", "-I" + os.path.join(ROOT, "oracle", "cvshim"), "-I" + os.path.join(REF, "include"), "-I" + REF,
           "-I" + os.path.join(ROOT, "include"), src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_dropin_signatures_are_the_reference_declarations():
    """Every method the drop-in defines is declared, with that exact signature, by the reference header it includes: a definition whose
    signature drifted would not compile (checked above); here: none of the drop-in files declares a class of its own."""
    for src in glob.glob(os.path.join(ROOT, "dropin", "*.cpp")):
        txt = open(src).read()
        assert not re.search(r"^\s*(class|struct)\s+(ORBMatcher|Optimizer|ORBextractor|System)\b", txt, flags=re.M), src


def test_integration_doc_names_the_renames_the_recipe_uses():
    """INTEGRATION.md tells a maintainer which reference bodies to rename away per file; oracle/Makefile builds libdropin.so with exactly those
    definitions. The two must not drift apart (a missing rename = a duplicate symbol at link time in the maintainer's tree)."""
    mk = open(os.path.join(ROOT, "oracle", "Makefile")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"^RENAME := (.*)$", mk, flags=re.M)
    assert m
    renames = re.findall(r"-D(\w+=\w+)", m.group(1))
    assert len(renames) >= 4
    for r in renames:
        assert r in doc, r
    # every drop-in source the recipe links is listed in the CMake snippet
    for src in sorted(glob.glob(os.path.join(ROOT, "dropin", "*_b200.cpp"))):
        assert os.path.basename(src) in doc, src
