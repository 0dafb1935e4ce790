"""bwamem-b200: a from-scratch B200-native BWA-MEM engine behind the gatk-bwamem-jni API.

The package holds the CUDA engine (csrc/ -> lib/libbwa_b200.so, C-ABI in include/bwamem_b200.h) and the host-side
mirror of the reference's Java classes (bwa.py).  The directory name contains '-', so load it by path
(see tests/common.py or __graft_entry__.py) rather than with a plain import statement."""
from . import bwa, shard
from .bwa import *  # noqa: F401,F403
