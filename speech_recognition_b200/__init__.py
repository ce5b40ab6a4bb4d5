"""kws-b200: B200-native keyword-spotting hot path of see--/speech_recognition.

Host code keeps the reference's call surface (AudioProcessor.get_data, load_model /
Model.predict, the submission CSV / uint8-memmap outputs); all arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``libkws.so`` (include/kws.h).
There is no CPU fallback.
"""
from .model_settings import prepare_model_settings  # noqa: F401
from .classes import (get_classes, get_int2label, get_label2int, prepare_words_list,  # noqa: F401
                      map_to_valid, map_to_wanted, class_map_32_to_12, AUDIO_NAMES)
from .engine import Engine, KwsError  # noqa: F401
from .model import Model, load_model, TTA_SHIPPED, TTA_8  # noqa: F401
from .audio_processor import AudioProcessor  # noqa: F401

__all__ = ["prepare_model_settings", "Engine", "KwsError", "Model", "load_model", "AudioProcessor",
           "TTA_SHIPPED", "TTA_8"]
