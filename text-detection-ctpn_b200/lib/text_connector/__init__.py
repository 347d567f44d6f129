"""Text-line stage: proposal graph, connectors, TextDetector."""
from .text_connect_cfg import Config            # noqa: F401
from .detectors import TextDetector             # noqa: F401
