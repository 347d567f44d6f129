#!/bin/bash
# Run ON THE GPU BOX: the per-layer conv timings / ablations behind profiles/r2_conv_ablation.txt (debug library switches).
python tools/time_conv.py --mode f16f8
CTPN_TC_STAGE_SMALL=0 python tools/time_conv.py --mode f16f8
python tools/time_conv.py --mode f16f8 --H 150 --W 225 --cin 256 --cout 256
CTPN_TC_STAGE_SMALL=0 python tools/time_conv.py --mode f16f8 --H 150 --W 225 --cin 256 --cout 256
python tools/time_conv.py --mode f16f8 --H 600 --W 900 --cin 64 --cout 64 --flags 3
CTPN_TC_DEBUG=4 python tools/time_conv.py --mode f16f8 --H 600 --W 900 --cin 64 --cout 64 --flags 3
CTPN_TC_DEBUG=16 python tools/time_conv.py --mode f16f8 --H 600 --W 900 --cin 64 --cout 64 --flags 3
python tools/time_conv.py --mode f16f8 --H 300 --W 450 --cin 128 --cout 128 --flags 3
python tools/time_conv.py --mode f16f8 --H 300 --W 450 --cin 64 --cout 128 --flags 1
python tools/time_bilstm.py
CTPN_LSTM_NC=2 python tools/time_bilstm.py
