from vad_b200.predictor import (VADFromScratchPredictor, VADPredictParameters,  # noqa: F401
                                merge_voice_activities)
