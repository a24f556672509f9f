"""bsmap_b200 -- B200-native drop-in for BSMAP 2.6's seed-and-extend bisulfite mapping hot path.

Host-side mirror of the reference's seam (main.cpp:49-114):

    RefSeq                  -> Index      (Run_ConvertBinseq + CreateIndex; device-resident)
    ReadClass               -> Reads      (CheckFile + LoadBatchReads; mmap + multi-threaded record cutter)
    SingleAlign / PairAlign -> Mapper     (ImportBatchReads + Do_Batch -> records -> SAM/BSP text)

All compute happens in libbsmap_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/bsmap_b200.h).  No CPU fallback exists.
"""
from .api import Index, Mapper, Meth, Reads, emit_pe, emit_se, make_params, meth_opts, pack_reads, pack_reads_2bit, sam_to_sorted_bam  # noqa: F401
from .lib import BsxError, Params, REC, PAIR_REC  # noqa: F401

__all__ = ["Index", "Mapper", "Reads", "Meth", "meth_opts", "emit_se", "emit_pe", "sam_to_sorted_bam", "make_params", "pack_reads", "pack_reads_2bit", "BsxError", "Params", "REC", "PAIR_REC"]
