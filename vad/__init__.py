"""Import-compatibility namespace: ``from vad.predictor import VADFromScratchPredictor`` etc.
resolve to the B200 implementations (drop-in for the reference's import surface on this path)."""
