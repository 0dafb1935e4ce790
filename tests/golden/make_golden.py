"""Regenerates tests/golden from the reference checkout (run in the build container only; /root/reference does not
exist on the GPU box).

rotavirus/   the reference's own test fixture, /root/reference/src/test/resources/ref.fa{,.amb,.ann,.bwt,.pac,.sa}
             (config #1 of BASELINE.json; pins the on-disk FM-index / SA / pac formats)
kat.json     the known-answer vectors asserted by
             /root/reference/src/test/java/org/broadinstitute/hellbender/utils/bwa/BwaMemIndexTest.java:45-127
"""
import json
import os
import shutil

SRC = "/root/reference/src/test/resources"
HERE = os.path.dirname(os.path.abspath(__file__))

KAT = {
    "source": "BwaMemIndexTest.java:45-127",
    "testSimple": {
        "reads": ["GGCTTTTAATGCTTTTCAGTGGTTGCTGCTCAAGATGGAGTCTACTCAGCAGATGGTAAGCTCTATTATT"],
        "expect": [[0, 70, 0, 70, "70M", 0, 0]],
    },
    "testMulti": {
        "reads": ["GGCTTTTAATGCTTTTCAGTGCTAGGTGCTCAAGATGGAGTCTACTCAGCAGATGGTAAGCTCTATTATT",
                  "AATAATAGAGCTTACCATCTGCTGAGTAGACTCCATCTTGAGCAGCAACCACTGAAAAGCATTAAAAGCC",
                  "AATACTTCTTTTGAAGCTGCAGTTGTTGCTGCCTTCAACATTAGAATTAATGGGTATTCAATATGATT"],
        "expect": [[0, 70, 0, 70, "70M", 3, 0], [0, 70, 0, 70, "70M", 0, 16], [70, 140, 0, 68, "32M2D36M", 2, 0]],
    },
    "testPair": {
        "reads": ["GGCTTTTAATGCTTTTCAGTGGTTGCTGCTCAAGATGGAGTCTACTCAGCAGATGGTAAGCTCTATTATT",
                  "TTGTTTTTAACACCAGAGTCATCCATCACATAATCAAATTTACTTTTAACTCTGGTAAATACTTCATTGT"],
        "stats_mode_1": [200, 10, 1, 600],
        "expect": {"0": [[0, 70, 0, 70, "70M", 0, 0x61, 140, 210], [140, 210, 0, 70, "70M", 0, 0x91, 0, -210]],
                   "1": [[0, 70, 0, 70, "70M", 0, 0x63, 140, 210], [140, 210, 0, 70, "70M", 0, 0x93, 0, -210]],
                   "2": [[0, 70, 0, 70, "70M", 0, 0x61, 140, 210], [140, 210, 0, 70, "70M", 0, 0x91, 0, -210]]},
    },
}

if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "rotavirus"), exist_ok=True)
    for f in os.listdir(SRC):
        if f.startswith("ref.fa"):
            shutil.copyfile(os.path.join(SRC, f), os.path.join(HERE, "rotavirus", f))
    with open(os.path.join(HERE, "kat.json"), "w") as fh:
        json.dump(KAT, fh, indent=1)
