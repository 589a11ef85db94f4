"""Golden vectors for the reference's own end-to-end fixture (tests/test_predict.py:12-30 of the
reference): the wav it predicts on + its test checkpoint.  Run once in the build container:

    python tests/golden/make_wav_golden.py

  * copies the reference TEST DATA file tests/data/WhenTheWeatherIsFine/When_the_Weather_Is_Fine_12_4.wav
    (a fixture, not source) next to this script;
  * extracts log-mel features with the repository's host FeatureExtractor (vad_b200.features -- librosa is
    absent here, so the features are this repo's NumPy restatement: "identical log-mel inputs" on both sides);
  * runs the UNMODIFIED reference SelfAttentiveVAD with the reference test checkpoint's weights through the
    literal per-item transcription of vad/predictor.py:169-258 (make_golden.predictor_loop_transcription);
  * stores the [L, 7] probabilities, the [L] mean (vad/predictor.py:95) and the checkpoint's config.
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VAD_REFERENCE", "/root/reference")
WAV = "When_the_Weather_Is_Fine_12_4.wav"

sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (puts the reference's `vad` package first on sys.path)

# the product package is reached through its real directory: `vad_b200` is an alias that would pull in the
# repository's `vad` compatibility package, which must not shadow the reference's here
import importlib.util  # noqa: E402


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    shutil.copyfile(os.path.join(REF, "tests/data/WhenTheWeatherIsFine", WAV), os.path.join(HERE, WAV))
    pkg = _load("vadb_pkg", os.path.join(ROOT, "voice-activity-detection_b200", "__init__.py"))
    from vadb_pkg.data_models import AudioData
    from vadb_pkg.features import FeatureExtractor
    ck = MG.load_sample_checkpoint()
    cfg = ck["config"]
    audio = AudioData.load(os.path.join(HERE, WAV))
    fe = FeatureExtractor(cfg["feature_extractor"])
    feat = fe.extract_with_postprocessing(audio)
    half = cfg["context_resolution"]["context_window_half_frames"]
    jump = cfg["context_resolution"]["context_window_jump_frames"]
    m = MG.ref_model({k: v.clone() for k, v in ck["state_dict"].items()})
    probs = MG.predictor_loop_transcription(m, feat, half, jump)
    np.savez_compressed(os.path.join(HERE, "weather_wav_golden.npz"), probs=probs.astype(np.float32),
                        mean=probs.mean(axis=1).astype(np.float32), n_samples=np.int64(len(audio.audio)),
                        config_json=np.array(json.dumps(cfg)))
    print("features", feat.shape, "probs", probs.shape, "mean P(speech)", float(probs.mean()),
          "frames > 0.5:", int((probs.mean(axis=1) > 0.5).sum()))


if __name__ == "__main__":
    main()
