"""The JVM-side binding: jni/vrec_jni.c must compile against include/vrec.h -- with a stand-in
jni.h, since the image has no JDK -- and define one stub for every @native method of Vrec.scala, each calling the
entry point it names.  The Scala drop-in classes must keep the reference's constructor and method signatures."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JNI = os.path.join(ROOT, "jni")
SCALA = os.path.join(JNI, "scala", "com", "github", "tashoyan", "recommender")


def test_jni_shim_type_checks_against_the_abi():
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
           "-I", os.path.join(ROOT, "tests", "jni_stub"), "-I", os.path.join(ROOT, "include"),
           os.path.join(JNI, "vrec_jni.c")]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr


def test_every_native_method_has_a_stub_that_calls_its_entry_point():
    scala = open(os.path.join(SCALA, "gpu", "Vrec.scala")).read()
    c = open(os.path.join(JNI, "vrec_jni.c")).read()
    header = open(os.path.join(ROOT, "include", "vrec.h")).read()
    natives = re.findall(r"@native def (\w+)\((?:.|\n)*?\)\s*:\s*\w+\s*//\s*(vrec_\w+)", scala)
    assert len(natives) >= 15
    for method, entry in natives:
        m = re.search(r"VREC_JNI\(\w+, %s\)\((?:.|\n)*?\n}\n" % method, c)
        assert m, f"no JNI stub for Vrec.{method}"
        assert entry + "(" in m.group(0), f"Vrec.{method} does not call {entry}"
        assert re.search(r"\b%s\(" % entry, header), f"{entry} is not declared in vrec.h"


def test_scala_drop_ins_keep_the_reference_signatures():
    knn = open(os.path.join(SCALA, "knn", "KnnRecommender.scala")).read()
    # knn/KnnRecommender.scala:9-16,22
    for arg in ("placeRatingVectors: DataFrame", "categoryRatingVectors: DataFrame", "placeRatings: DataFrame",
                "placeWeight: Double", "categoryWeight: Double", "kNearest: Int"):
        assert arg in knn
    assert "def makeRecommendations(personId: Long): DataFrame" in knn
    assert 'IllegalArgumentException(s"No such person: $personId")' in knn
    sg = open(os.path.join(SCALA, "stochastic", "StochasticRecommender.scala")).read()
    # stochastic/StochasticRecommender.scala:28-32,66
    for arg in ("stochasticEdges: DataFrame", "epsilon: Double", "maxIterations: Int"):
        assert arg in sg
    assert "def makeRecommendations(vertexId: Long): DataFrame" in sg
    assert 'IllegalArgumentException(s"No such vertex in the graph: $vertexId")' in sg
