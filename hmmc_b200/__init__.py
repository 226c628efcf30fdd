"""hmmc_b200: B200 (sm_100a) implementation of the HMMC hierarchical-matching contrastive
head.  Python keeps the reference's interface (modeling / metrics / retrieval); the
arithmetic lives in libhmmc_head.so (include/hmmc_head.h)."""
__version__ = "0.1.0"
