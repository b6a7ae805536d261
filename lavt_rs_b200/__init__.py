"""lavt_rs_b200 -- B200-native (sm_100a) implementation of the LAVT-RS language-aware Swin hot path.

    from lavt_rs_b200.lib import segmentation
    model = segmentation.lavt_video(pretrained="", args=args).cuda().eval()
    logits = model(frames, token_ids, attention_mask)        # same call as the reference

Kernels: ``csrc/*.cu`` behind the C ABI in ``include/lavt_b200.h`` (built by ``python -m lavt_rs_b200.build``).
"""
__version__ = "0.1.0"
