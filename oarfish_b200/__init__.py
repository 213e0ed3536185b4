"""oarfish_b200 -- B200-native EM / bootstrap engine behind oarfish's em::em,
em::em_par and em::bootstrap (reference: COMBINE-lab/oarfish src/em.rs).

Layout: csrc/ holds the sm_100a kernels and the C ABI (include/oarfish_em.h);
em.py mirrors the reference's interface; engine.py wraps the ABI handle.
"""
from .em import (ALN_INFO_DTYPE, AlignmentFilters, EMInfo, InMemoryAlignmentStore, TranscriptInfo, bootstrap, em,
                 em_par)
from .engine import DeviceStore, EMResult, MultiStore, device_count, em_batched_multi

__all__ = ["ALN_INFO_DTYPE", "AlignmentFilters", "EMInfo", "InMemoryAlignmentStore", "TranscriptInfo", "bootstrap",
           "em", "em_par", "DeviceStore", "EMResult", "MultiStore", "device_count", "em_batched_multi"]
