"""qdiff — B200-native mirror of the reference's quantization plugin surface
(ViDiT-Q/quant_utils/qdiff): same import paths, class names, attributes and quant_param_dict
schema, so the reference's callers (examples/Wan2.1/{quant_generate,ptq_wanx,get_calib_data_wanx}.py,
wan/quant_wanx.py:13-17) run unchanged once this directory precedes the reference's on sys.path.
All arithmetic is executed by libb200q (CUDA, sm_100a) through `b200q`; there is no torch
fake-quant fallback."""
