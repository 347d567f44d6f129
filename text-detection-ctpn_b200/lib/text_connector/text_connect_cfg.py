"""Constants of the text-line stage under the names the reference's callers use
(lib/text_connector/text_connect_cfg.py:2-12; read as `Config.X` / `TextLineCfg.X`, mutable class attributes)."""


class Config:
    SCALE = 600
    MAX_SCALE = 1200
    TEXT_PROPOSALS_WIDTH = 16
    MIN_NUM_PROPOSALS = 2
    MIN_RATIO = 0.5
    LINE_MIN_SCORE = 0.9
    MAX_HORIZONTAL_GAP = 50
    TEXT_PROPOSALS_MIN_SCORE = 0.7
    TEXT_PROPOSALS_NMS_THRESH = 0.2
    MIN_V_OVERLAPS = 0.7
    MIN_SIZE_SIM = 0.7


def native_cfg():
    """The constants in the order ctpn_text_*_host take them (include/ctpn_b200.h), read at call time."""
    c = Config
    return (c.TEXT_PROPOSALS_MIN_SCORE, c.TEXT_PROPOSALS_NMS_THRESH, c.MAX_HORIZONTAL_GAP, c.MIN_V_OVERLAPS, c.MIN_SIZE_SIM,
            c.MIN_RATIO, c.LINE_MIN_SCORE, c.TEXT_PROPOSALS_WIDTH, c.MIN_NUM_PROPOSALS)
