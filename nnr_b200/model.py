"""Drop-in ``Model`` (reference model.py:10-133) for the CNE / SUE family with the dot-product predictor."""
import torch.nn as nn

from . import engine, newsEncoders, userEncoders, variantEncoders


class Model(nn.Module):
    def __init__(self, config):
        super().__init__()
        if config.news_encoder == 'CNE':
            self.news_encoder = newsEncoders.CNE(config)
        elif config.news_encoder == 'CNE_wo_CA':
            self.news_encoder = variantEncoders.CNE_wo_CA(config)
        elif config.news_encoder in ('CNE_wo_CS', 'CNE_Title', 'CNE_Content'):
            self.news_encoder = getattr(variantEncoders, config.news_encoder)(config)
        else:
            raise Exception(config.news_encoder + ' is outside the nnr_b200 hot path (CNE and its ablations)')
        if config.user_encoder == 'SUE':
            self.user_encoder = userEncoders.SUE(self.news_encoder, config)
        elif config.user_encoder == 'SUE_wo_HCA':
            self.user_encoder = variantEncoders.SUE_wo_HCA(self.news_encoder, config)
        elif config.user_encoder == 'SUE_wo_GCN':
            self.user_encoder = variantEncoders.SUE_wo_GCN(self.news_encoder, config)
        else:
            raise Exception(config.user_encoder + ' is outside the nnr_b200 hot path (SUE, SUE_wo_HCA, SUE_wo_GCN)')
        self.model_name = config.news_encoder + '-' + config.user_encoder
        self.news_embedding_dim = self.news_encoder.news_embedding_dim
        self.dropout = nn.Dropout(p=config.dropout_rate)
        self.use_user_embedding = False
        if config.click_predictor != 'dot_product':
            raise Exception('nnr_b200 implements the dot_product click predictor (reference default, config.py:76)')
        self.click_predictor = config.click_predictor

    def initialize(self):
        self.news_encoder.initialize()
        self.user_encoder.initialize()
        engine.weights_changed()

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        engine.weights_changed()       # cached operand planes of the weight matrices are stale
        return out

    def forward(self, user_ID, user_category, user_subCategory, user_title_text, user_title_mask, user_title_entity,
                user_content_text, user_content_mask, user_content_entity, user_history_mask, user_history_graph,
                user_history_category_mask, user_history_category_indices, news_category, news_subCategory,
                news_title_text, news_title_mask, news_title_entity, news_content_text, news_content_mask,
                news_content_entity):
        # One CNE kernel schedule for both reference calls (model.py:123 candidates, userEncoders.py:76-78 history);
        # each keeps its own sort-rank pairing domain, so results equal two separate calls.
        news_representation, history_embedding = self.news_encoder.encode_calls([
            (news_title_text, news_title_mask, news_content_text, news_content_mask, news_category, news_subCategory),
            (user_title_text, user_title_mask, user_content_text, user_content_mask, user_category, user_subCategory)])
        user_representation = self.user_encoder.encode_user(history_embedding, user_history_graph, user_history_category_mask,
                                                            user_history_category_indices, news_representation)
        return engine.RowDot.apply(user_representation, news_representation)     # model.py:127
