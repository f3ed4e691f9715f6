"""Host-side mirror of the reference's output contract for the index path:
`Distance::try_from` (crates/vector-store/src/distance.rs:58-105) and
`SimilarityScore::from(Distance)` (similarity.rs:26-37).  Every hit the engine returns is passed
through `Distance.try_from`, exactly like vs_index/usearch.rs:219."""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass

import numpy as np


class SpaceType(enum.Enum):
    Euclidean = "EUCLIDEAN"
    Cosine = "COSINE"
    DotProduct = "DOT_PRODUCT"
    Hamming = "HAMMING"


@dataclass(frozen=True)
class Distance:
    space: SpaceType
    value: float
    dimensions: int | None = None

    @staticmethod
    def try_from(value: float, space: SpaceType, dimensions: int | None = None) -> "Distance":
        v = float(np.float32(value))
        if space is SpaceType.Cosine:
            if not (0.0 <= v <= 2.0):
                raise ValueError("Cosine distance must be in range [0.0, 2.0]")
        elif space is SpaceType.Euclidean:
            if not (v >= 0.0):
                raise ValueError("Euclidean distance must be >= 0.0")
        elif space is SpaceType.DotProduct:
            if math.isnan(v):
                raise ValueError("Dot Product distance must be a valid number, got NaN")
        else:
            if not (v >= 0.0):
                raise ValueError("Hamming distance must be >= 0.0")
            if not math.isfinite(v):
                raise ValueError("Hamming distance must be a finite number")
            if v != math.floor(v):
                raise ValueError("Hamming distance must be an integer value")
            if dimensions is None:
                raise ValueError("Dimensions must be provided for Hamming distance")
            if v > float(dimensions):
                raise ValueError("Hamming distance cannot be greater than the number of dimensions")
        return Distance(space, v, dimensions)

    def __float__(self) -> float:
        return self.value


@dataclass(frozen=True)
class SimilarityScore:
    value: float

    @staticmethod
    def from_distance(d: Distance) -> "SimilarityScore":
        x = np.float32(d.value)
        if d.space in (SpaceType.Cosine, SpaceType.DotProduct):
            s = (np.float32(2.0) - x) / np.float32(2.0)
        elif d.space is SpaceType.Euclidean:
            s = np.float32(1.0) / (np.float32(1.0) + x)
        else:
            s = np.float32(1.0) - x / np.float32(d.dimensions)
        return SimilarityScore(float(np.float32(s)))
