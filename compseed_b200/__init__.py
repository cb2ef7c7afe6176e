"""compseed_b200: B200-native (sm_100a) SMEM seeding path of CompSeed / BWA-MEM behind a C-ABI.

Only the seeding hot path lives here: csrc/ holds the CUDA kernels and the
C-ABI, seeding.py the ctypes mirror of the reference's interface, synth.py the synthetic inputs.
"""
from .seeding import (FMIndex, SeedContext, SeedOpt, SeedResult, CompSeedError, IndexConfig, CtxConfig, ChainOpt, ChainResult, seed_reads, device_count,
                      probe_random_gather, flush_l2, load_library, host_register, host_unregister, pack_reads, pack_reads_host, pack_reads_host64, packed_words, MultiSeeder, replicate_index,
                      BswExtender, BswOpt, bwa_fill_scmat)

__all__ = ["FMIndex", "SeedContext", "SeedOpt", "SeedResult", "CompSeedError", "IndexConfig", "CtxConfig", "ChainOpt", "ChainResult", "seed_reads", "device_count",
           "probe_random_gather", "flush_l2", "load_library", "host_register", "host_unregister", "pack_reads", "pack_reads_host", "pack_reads_host64", "packed_words", "MultiSeeder", "replicate_index",
           "BswExtender", "BswOpt", "bwa_fill_scmat"]
