from .configuration_bert import BertQuantizedConfig
from .modeling_bert import (BertQuantizedAttention, BertQuantizedEncoder, BertQuantizedForMaskedLM,
                            BertQuantizedForQuestionAnswering, BertQuantizedForSequenceClassification,
                            BertQuantizedForTokenClassification, BertQuantizedIntermediate, BertQuantizedLayer,
                            BertQuantizedModel, BertQuantizedOutput, BertQuantizedSelfAttention, BertQuantizedSelfOutput)
from .quant_config_bert import parse_bert_quantized_config
from .profiler_bert import profile_bert_quantized
