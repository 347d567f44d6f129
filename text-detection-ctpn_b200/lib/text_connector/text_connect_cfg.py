"""Constants of the text-line stage, under the names the reference's callers use
(lib/text_connector/text_connect_cfg.py; read as `Config.X` / `TextLineCfg.X`)."""

_IMAGE = dict(
    SCALE=600,                     # resize_im: short side
    MAX_SCALE=1200,                # resize_im: cap of the long side
)
_PROPOSALS = dict(
    TEXT_PROPOSALS_WIDTH=16,       # anchor width (px)
    TEXT_PROPOSALS_MIN_SCORE=0.7,  # proposals entering the connector
    TEXT_PROPOSALS_NMS_THRESH=0.2,
)
_GRAPH = dict(
    MAX_HORIZONTAL_GAP=50,         # px searched to the left / right for a neighbour
    MIN_V_OVERLAPS=0.7,            # vertical overlap of neighbours
    MIN_SIZE_SIM=0.7,              # height similarity of neighbours
)
_LINES = dict(
    MIN_NUM_PROPOSALS=2,
    MIN_RATIO=0.5,                 # width / height of a kept line
    LINE_MIN_SCORE=0.9,
)

Config = type("Config", (), {**_IMAGE, **_PROPOSALS, **_GRAPH, **_LINES,
                             "__doc__": "Text-line construction constants (mutable class attributes, as in the reference)."})
