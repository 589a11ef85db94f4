"""I/O data models of the predict/evaluate path, schema-compatible with the reference
(vad/data_models/audio_data.py, voice_activity.py, vad_data.py, vad/util/time_utils.py) so the
CLI output (VoiceActivity JSON v0.3) is a drop-in.  Pure host-side bookkeeping.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from datetime import datetime, timedelta
from enum import Enum
from pathlib import Path
from typing import List, Optional

import numpy as np

STANDARD_SAMPLE_RATE = 16000


# ---- time formatting (vad/util/time_utils.py:6-38) ----
def parse_timecode_to_timedelta(timecode: str) -> timedelta:
    return datetime.strptime(timecode, "%H:%M:%S.%f") - datetime(year=1900, month=1, day=1)


def format_timedelta_to_timecode(t: timedelta) -> str:
    total = int(t.total_seconds())
    ms = round(t.microseconds / 1000)
    return f"{total // 3600:02d}:{total % 3600 // 60:02d}:{total % 60:02d}.{ms:03d}"


def format_timedelta_to_milliseconds(t: timedelta) -> int:
    return int(t.total_seconds() * 1000)


# ---- audio (vad/data_models/audio_data.py:12-34) ----
@dataclass
class AudioData:
    audio: np.ndarray      # 1-D float32 samples
    sample_rate: int
    duration: timedelta

    @classmethod
    def load(cls, path: Path):
        path = Path(path)
        if path.suffix == ".pcm":
            audio = np.fromfile(str(path), dtype=np.int16).astype(np.single) / 32768
        else:
            audio, sample_rate = _read_wav(path)
            if sample_rate != STANDARD_SAMPLE_RATE:
                # reference: librosa.resample(..., res_type="kaiser_fast") (resampy, absent here);
                # polyphase resampling is the stand-in -- parity unpinned for non-16 kHz input
                from math import gcd
                from scipy.signal import resample_poly
                g = gcd(int(sample_rate), STANDARD_SAMPLE_RATE)
                audio = resample_poly(audio, STANDARD_SAMPLE_RATE // g, int(sample_rate) // g).astype(np.single)
        return cls(audio=audio, sample_rate=STANDARD_SAMPLE_RATE,
                   duration=timedelta(seconds=len(audio) / STANDARD_SAMPLE_RATE))


def _read_wav(path: Path):
    """soundfile.read(path, dtype=float32, always_2d=True).mean(axis=1) for PCM/float WAV."""
    from scipy.io import wavfile
    sample_rate, data = wavfile.read(str(path))
    if data.dtype == np.int16:
        audio = data.astype(np.single) / 32768
    elif data.dtype == np.int32:
        audio = (data.astype(np.float64) / 2147483648).astype(np.single)
    elif data.dtype == np.uint8:
        audio = (data.astype(np.single) - 128) / 128
    else:
        audio = data.astype(np.single)
    if audio.ndim == 2:
        audio = audio.mean(axis=1)
    return audio, int(sample_rate)


# ---- voice activity (vad/data_models/voice_activity.py) ----
class VoiceActivityVersion(Enum):
    v01 = "v0.1"
    v02 = "v0.2"
    v03 = "v0.3"


@dataclass
class Activity:
    start: timedelta
    end: timedelta


@dataclass
class VoiceActivity:
    duration: timedelta
    activities: List[Activity]
    probs_sample_rate: Optional[int]
    probs: Optional[List[float]]

    @classmethod
    def load(cls, path: Path):
        with Path(path).open() as f:
            return cls.from_json(json.load(f))

    @classmethod
    def from_json(cls, data: dict):
        version = data["version"]
        common = dict(probs_sample_rate=data.get("probs_sample_rate"), probs=data.get("probs"))
        if version == "v0.3":
            return cls(duration=parse_timecode_to_timedelta(data["duration"]),
                       activities=[Activity(parse_timecode_to_timedelta(a["start"]),
                                            parse_timecode_to_timedelta(a["end"]))
                                   for a in data["activities"]], **common)
        if version in ("v0.1", "v0.2"):
            if version == "v0.2" and data.get("time_format") == "millisecond":
                return cls(duration=timedelta(milliseconds=data["duration"]),
                           activities=[Activity(timedelta(milliseconds=b["start_time"]),
                                                timedelta(milliseconds=b["end_time"]))
                                       for b in data["voice_activity"]], **common)
            if version == "v0.2" and data.get("time_format") != "timecode":
                raise NotImplementedError
            return cls(duration=parse_timecode_to_timedelta(data["duration"]),
                       activities=[Activity(parse_timecode_to_timedelta(b["start_time"]),
                                            parse_timecode_to_timedelta(b["end_time"]))
                                   for b in data["voice_activity"]], **common)
        raise NotImplementedError(version)

    def to_json(self, version: VoiceActivityVersion = VoiceActivityVersion.v03) -> dict:
        if version == VoiceActivityVersion.v03:
            return {"version": "v0.3",
                    "duration": format_timedelta_to_timecode(self.duration),
                    "activities": [{"start": format_timedelta_to_timecode(a.start),
                                    "end": format_timedelta_to_timecode(a.end)}
                                   for a in self.activities],
                    "probs_sample_rate": self.probs_sample_rate, "probs": self.probs}
        blocks = [{"start_time": format_timedelta_to_timecode(a.start),
                   "end_time": format_timedelta_to_timecode(a.end)} for a in self.activities]
        out = {"version": version.value, "duration": format_timedelta_to_timecode(self.duration)}
        if version == VoiceActivityVersion.v02:
            out["time_format"] = "timecode"
        out.update({"voice_activity": blocks, "probs_sample_rate": self.probs_sample_rate,
                    "probs": self.probs})
        return out

    def save(self, path: Path, version: VoiceActivityVersion = VoiceActivityVersion.v03):
        with Path(path).open("w") as f:
            json.dump(self.to_json(version), f, ensure_ascii=False, indent=4)

    def to_labels(self, sample_rate: int) -> np.ndarray:
        labels = np.zeros(int(self.duration.total_seconds() * sample_rate), dtype=np.int64)
        for a in self.activities:
            labels[int(a.start.total_seconds() * sample_rate):
                   int(a.end.total_seconds() * sample_rate)] = 1
        return labels


# ---- evaluation lists (vad/data_models/vad_data.py) ----
@dataclass
class VADDataPair:
    audio_path: Path
    voice_activity_path: Path


@dataclass
class VADDataList:
    pairs: List[VADDataPair]

    @classmethod
    def load(cls, path: Path):
        pairs = []
        with Path(path).open() as f:
            for line in f:
                if line.strip():
                    d = json.loads(line)
                    pairs.append(VADDataPair(Path(d["audio_path"]), Path(d["voice_activity_path"])))
        return cls(pairs=pairs)
