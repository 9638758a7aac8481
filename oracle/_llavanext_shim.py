"""transformers-5.x adapter around the reference's LlavaNextForRL (models/LlavaNext/__init__.py).  Imported via ref_shim."""
from transformers.modeling_outputs import CausalLMOutputWithPast
from vlrlhf.models.LlavaNext import LlavaNextForRL


class LlavaNextShim(LlavaNextForRL):
    @property
    def vision_tower(self):
        return self.model.vision_tower

    @property
    def multi_modal_projector(self):
        return self.model.multi_modal_projector

    @property
    def image_newline(self):
        return self.model.image_newline

    @property
    def padding_side(self):
        return "right"

    def pack_image_features(self, image_features, image_sizes, image_newline=None):
        feats, lens = self.model.pack_image_features(image_features, image_sizes, vision_feature_select_strategy="default",
                                                     image_newline=image_newline)
        if isinstance(feats, (list, tuple)):  # transformers 4.41 returned the concatenation (modeling_llava_next.py)
            import torch
            feats = torch.cat(list(feats), dim=0)
        if not hasattr(lens, "sum") or isinstance(lens, (list, tuple)):
            import torch
            lens = torch.tensor(list(lens), dtype=torch.long)
        return feats, lens

    def language_model(self, **kw):
        kw.pop("return_dict", None)
        kw.pop("output_attentions", None)
        o = self.model.language_model(**kw)
        return CausalLMOutputWithPast(logits=self.lm_head(o.last_hidden_state).float(),
                                      past_key_values=o.past_key_values, hidden_states=o.hidden_states)
