"""adyolo_b200 — B200-native drop-in for AD-YOLO's data-parallel hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``include/adyolo_b200.h``
(``lib/libadyolo_b200.so``, built in-tree by ``__graft_entry__.build()``).  There is no CPU or
PyTorch fallback: every entry point raises if the library or a CUDA device is missing.

Mirrors (same names / argument meaning) of the reference call surface, SURVEY.md §8(b):

* ``features.FeatureLabelProcessor``, ``features.audio2stft / stft2melscale / stft2iv``
  (src/datasets.py:187-292, src/utils/utility.py:142-215) + batched ``features_batched``
* ``labels.get_yolo_label`` / ``labels.collate_fn`` (src/datasets.py:457-482, 164-184)
* ``loss.ADYOLOloss`` / ``loss.WrapperCriterion`` (src/models/loss.py:156-251, src/wrapper.py:63-88)
* ``scaler.preprocess_scaler`` (src/preprocess.py:87-130) with a multi-GPU all-reduce
* ``augment.SpecAug`` (src/utils/augmentations.py:6-33); rotation is fused into features / labels
"""
from . import _lib  # noqa: F401
from .features import (FeatureLabelProcessor, FrontEndModule, audio2stft, stft2melscale, stft2iv,  # noqa: F401
                       features_batched, mel_filterbank)
from .labels import DeviceRows, get_yolo_label, collate_fn, label_rows_batched  # noqa: F401
from .loss import ADYOLOloss, WrapperCriterion, adyolo_assign  # noqa: F401
from .scaler import ScalerAccumulator, preprocess_scaler  # noqa: F401
from .pipeline import HostBatchPipeline, bind_host_to_device  # noqa: F401
from .postprocess import LabelPostProcessor, yolo_post_batched  # noqa: F401
from .augment import RotationAug, SpecAug  # noqa: F401
from .data import (ResidentClips, EpochSampler, RawAudioDataset, collate_raw, chunk_plan,  # noqa: F401
                   features_batched_views, load_wav2npy, load_csv2dict)

__version__ = "0.1.0"
