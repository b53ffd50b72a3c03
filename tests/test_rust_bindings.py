"""CPU: the Rust `-sys` binding (bindings/rust/src/sys.rs) is generated from include/takzero_b200.h and must stay in
step with it and with the shared library.  There is no Rust toolchain in the image, so this is a consistency check
of an UNVERIFIED binding, not a build."""
import os
import re
import subprocess
import sys

from takzero_b200 import build as tz_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generated_binding_is_up_to_date():
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_bindings.py"), "--check"]).returncode
    assert rc == 0, "run `python tools/gen_rust_bindings.py` after changing include/takzero_b200.h"


def test_binding_declares_every_exported_symbol_and_the_wrapper_uses_declared_ones():
    out = subprocess.check_output(["nm", "-D", "--defined-only", tz_build.build()], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    sysrs = open(os.path.join(ROOT, "bindings", "rust", "src", "sys.rs")).read()
    declared = set(re.findall(r"pub fn (tz_\w+)\(", sysrs))
    assert declared == set(exported)
    wrapper = open(os.path.join(ROOT, "bindings", "rust", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(tz_\w+)\(", wrapper))
    assert used and used <= declared
    # struct layouts the wrapper relies on
    assert "pub stack: [u64; TZ_MAX_SQ]," in sysrs and "pub tree_batch: c_int," in sysrs
