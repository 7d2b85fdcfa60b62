"""MedSAM wrapper with the reference's constructor, freezing policy and ``forward(image, box)`` signature
(reference ``segment_anything/network.py``)."""
import weakref

import torch
import torch.nn as nn

from segment_anything import sam_model_registry  # noqa: F401  (the reference module re-exports it, network.py:3)

from .modeling.common import require_inference


class MedSAM(nn.Module):
    def __init__(self, image_encoder, mask_decoder, prompt_encoder):
        super().__init__()
        self.image_encoder = image_encoder
        self.mask_decoder = mask_decoder
        self.prompt_encoder = prompt_encoder
        # same freezing policy as the reference (network.py:19-34): only the HFC adaptor and patch embed of the
        # encoder are trainable; prompt encoder and decoder are trainable
        for name, param in self.image_encoder.named_parameters():
            param.requires_grad = ("hfc_embed" in name) or ("hfc_attn" in name) or ("patch_embed" in name)
        for param in self.prompt_encoder.parameters():
            param.requires_grad = True
        for param in self.mask_decoder.parameters():
            param.requires_grad = True

    def fft(self, img, rate=0.125):
        """High-frequency image |g - lowpass(g)| [B,1,1024,1024] fp32 (reference network.py:36-57), computed as two
        DFT-operator GEMMs instead of fft2/ifft2.  ``rate`` other than 0.125 is not supported."""
        if rate != 0.125:
            raise NotImplementedError("the low-pass operator is built for the reference's rate=0.125")
        x = img.tensors if hasattr(img, "tensors") else img
        require_inference(self, x)
        xc = x.contiguous().float()
        eng = self.image_encoder.engine()
        a_patch, a_hfc, hfc_img = eng.highpass(xc, want_image=True)
        if xc is x:  # (a converted temporary could be freed and its address reused: never key a cache on it)
            hfc_img._wm_rows = (weakref.ref(x), x._version, hfc_img._version, eng, eng.gen_rows, a_patch, a_hfc)
        return hfc_img

    def forward(self, image, box):
        """image: NestedTensor (``.tensors`` [B,3,1024,1024] fp32 on the GPU); ``box`` is accepted and ignored like
        in the reference (network.py:69-78).  Returns {'pred_logits': [B,Q,8], 'pred_boxes': [B,Q,4]} fp32."""
        x = image.tensors if hasattr(image, "tensors") else image
        # = image_encoder(x, self.fft(image)) (network.py:80-81) without materialising the fp32 high-pass image
        image_embedding = self.image_encoder.forward_tiles(x)
        return self.mask_decoder(
            image_embeddings=image_embedding,
            image_pe=self.prompt_encoder.get_dense_pe(),
            sparse_prompt_embeddings=None,
            dense_prompt_embeddings=None,
            multimask_output=False,
            hfc_embed=None,
        )
