from .aux_functions import add_layer_norm, gen_encoder_output_proposals, prepare_value, topk_proposals
from .msda_function import MSDeformAttnFunction, MSDeformAttnFusedFunction

__all__ = ["MSDeformAttnFunction", "MSDeformAttnFusedFunction", "prepare_value", "gen_encoder_output_proposals",
           "topk_proposals"]
