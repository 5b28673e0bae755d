"""B200-native drop-in for the speaker-embedding extraction path of mvector 1.1.1
(yeyupiaoling/VoiceprintRecognition-Pytorch): same ``mvector.predict.MVectorPredictor`` / ``configs/*.yml`` surface,
hand-written sm_100a CUDA underneath (libvpb200.so, include/vpb200.h).  No CPU fallback."""
__version__ = '1.1.1+b200.r1'
