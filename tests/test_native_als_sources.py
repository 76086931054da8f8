"""The native-als module (Scala + JNI sources a PredictionIO checkout compiles) cannot be built in a
container without a JDK, but it must not rot: the JNI shim is type-checked against include/pio_als.h with a stand-in
jni.h, and every `@native` method of the Scala object must have its C counterpart (and vice versa)."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
JNI_C = ROOT / "native-als" / "src" / "main" / "c" / "pio_als_jni.c"
SCALA = ROOT / "native-als" / "src" / "main" / "scala" / "org" / "apache" / "predictionio" / "nativeals" / "NativeALS.scala"


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_jni_shim_type_checks_against_the_c_abi():
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                        "-I", str(ROOT / "tests" / "jni_mock"), "-I", str(ROOT / "include"), str(JNI_C)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_scala_native_has_a_jni_function():
    scala = set(re.findall(r"@native def (\w+)\(", SCALA.read_text()))
    c = set(re.findall(r"JNAME\((\w+)\)\(JNIEnv", JNI_C.read_text()))
    assert scala and scala == c, (sorted(scala - c), sorted(c - scala))


def test_jni_shim_calls_only_declared_abi_functions():
    declared = set(re.findall(r"PIO_API\s+[\w\s\*]+?\b(pio_\w+)\(", (ROOT / "include" / "pio_als.h").read_text()))
    used = set(re.findall(r"\b(pio_(?:als|nb)_\w+)\(", JNI_C.read_text()))
    assert used <= declared, sorted(used - declared)
