"""B200-native YOLO detection post-processing (decode + NMS) behind the reference's own call signatures.

Public surface (mirrors ultralytics/utils/nms.py and ultralytics/nn/modules/head.py of Chriz122/ultralytics_pro):
    non_max_suppression, TorchNMS            - ultralytics_pro_b200.nms
    detect_inference, decode_head            - ultralytics_pro_b200.head
    postprocess_from_head                    - fused decode + NMS
    install / uninstall                      - ultralytics_pro_b200.patch (rebinds the reference's symbols)
"""
__version__ = "0.1.0"

__all__ = ["non_max_suppression", "TorchNMS", "detect_inference", "decode_head", "postprocess_from_head"]


def __getattr__(name):
    if name in ("non_max_suppression", "TorchNMS"):
        from . import nms

        return getattr(nms, name)
    if name in ("detect_inference", "decode_head", "postprocess_from_head"):
        from . import head

        return getattr(head, name)
    raise AttributeError(name)
