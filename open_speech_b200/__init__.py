"""open_speech_b200 -- B200-native (sm_100a CUDA) audio-signal hot path of open-speech.

Keeps the reference's Python signatures as a drop-in boundary:

    reference module                  this package
    src.audio.preprocessing      ->   open_speech_b200.audio.preprocessing
    src.audio.postprocessing     ->   open_speech_b200.audio.postprocessing
    src.realtime.audio_buffer    ->   open_speech_b200.realtime.audio_buffer
    src.streaming.resample_pcm16 ->   open_speech_b200.streaming.resample_pcm16
    src.vad.silero               ->   open_speech_b200.vad.silero
    src.effects.chain            ->   open_speech_b200.effects.chain
    src.tts.voices / pipeline    ->   open_speech_b200.tts.voices / pipeline
    faster_whisper FeatureExtractor-> open_speech_b200.features.B200FeatureExtractor

Every numeric call goes through libosb200.so (include/osb200.h).  There is no CPU or
PyTorch fallback: a missing library or GPU raises RuntimeError.
"""
__version__ = "0.1.0"


def install_dropin() -> None:
    """Alias the reference's module paths onto this package."""
    from .dropin import install

    install()
