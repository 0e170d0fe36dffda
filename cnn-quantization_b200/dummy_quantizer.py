"""Identity quantizer (reference: pytorch_quantizer/quantization/qtypes/dummy_quantizer.py:1-7)."""


class DummyQuantizer:
    def __call__(self, tensor, id=None, tag="", stat_id=None, override_att=None, weight_correction=None):
        return tensor

    def __repr__(self):
        return "DummyQuantizer - fp32"
