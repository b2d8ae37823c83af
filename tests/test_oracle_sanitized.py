"""AddressSanitizer + UndefinedBehaviorSanitizer over the native CPU code the parity claims rest on (SURVEY.md §5:
"CPU oracle under -fsanitize=address,undefined"): the C oracle (`make -C oracle sanitize`) and the host build of the
product's sampler control logic (csrc/nuts_core.h, tests/emul) run the density, sampler-control, random-design and
Pathfinder tests again in a child interpreter with libasan preloaded; any report fails the test."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_and_control_logic_are_clean_under_asan_ubsan():
    asan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not installed")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "sanitize"])
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0",
               UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1", CSPLOTCH_EMUL_SAN="1",
               CSPLOTCH_ORACLE_LIB=os.path.join(ROOT, "oracle", "_build", "libcsplotch_oracle_san.so"))
    files = ["tests/test_oracle_density.py", "tests/test_nuts_core_vs_oracle.py", "tests/test_oracle_random_designs.py",
             "tests/test_pathfinder.py", "tests/test_vinit_host.py"]
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider"] + files, cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=1500)
    out = r.stdout + r.stderr
    assert "AddressSanitizer" not in out and "runtime error:" not in out, out[-4000:]
    assert r.returncode == 0, out[-4000:]
    assert " passed" in out
