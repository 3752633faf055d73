from .modules import (SinusoidalPositionalEmbedding, RPEMultiHeadAttention, RPEAttentionLayer, RPETransformerLayer,
                      MultiHeadAttention, AttentionLayer, TransformerLayer, AttentionOutput, RPEConditionalTransformer)
