"""transformers-5.x adapter around the reference's LlavaForRL (must live in a real .py file:
PreTrainedModel.__init__ inspects the class's source file).  Imported via ref_shim only."""
from transformers.modeling_outputs import CausalLMOutputWithPast
from vlrlhf.models.Llava import LlavaForRL


class LlavaShim(LlavaForRL):
    @property
    def vision_tower(self):
        return self.model.vision_tower

    @property
    def multi_modal_projector(self):
        return self.model.multi_modal_projector

    @property
    def pad_token_id(self):
        return self.config.pad_token_id if self.config.pad_token_id is not None else -1

    def language_model(self, **kw):
        kw.pop("return_dict", None)
        kw.pop("output_attentions", None)
        o = self.model.language_model(**kw)
        # .float(): transformers 4.41 LlamaForCausalLM upcasts logits (SURVEY.md §8c caveat i)
        return CausalLMOutputWithPast(logits=self.lm_head(o.last_hidden_state).float(),
                                      past_key_values=o.past_key_values, hidden_states=o.hidden_states)
