"""LSTM language models.

Parity: ``fedml_api/model/nlp/rnn.py:4-33`` (RNN_OriginalFedAvg: Embedding(90,8)
→ 2-layer LSTM(256, batch_first) → Linear(256,90) on the LAST step; 822 570
params; on CUDA the embedding + 2-layer LSTM run in the fused persistent kernel ``csrc/lstm_tc.cu`` — state-dict
keys stay those of ``nn.Embedding`` / ``nn.LSTM``) and ``:36-66`` (RNN_StackOverFlow: Embedding(10004,96) → LSTM(670) →
fc 96 → fc 10004; 4 053 428 params).  ``per_position=True`` is the optional
TFF-style per-position head.
"""
from __future__ import annotations

from torch import nn

from ..ops.linear import TcLinear


class RNN_OriginalFedAvg(nn.Module):
    def __init__(self, embedding_dim: int = 8, vocab_size: int = 90, hidden_size: int = 256,
                 per_position: bool = False):
        super().__init__()
        self.per_position = per_position
        self.embeddings = nn.Embedding(num_embeddings=vocab_size, embedding_dim=embedding_dim, padding_idx=0)
        self.lstm = nn.LSTM(input_size=embedding_dim, hidden_size=hidden_size, num_layers=2, batch_first=True)
        self.fc = TcLinear(hidden_size, vocab_size)

    def forward(self, input_seq):
        from ..ops import lstm as fused
        if fused.eligible(input_seq, self.embeddings.weight, self.lstm):
            # persistent cluster-resident tcgen05 kernel: the whole sequence, both layers, ONE launch (csrc/lstm_tc.cu)
            if self.per_position:
                out = fused.lstm2_embed_forward(input_seq, self.embeddings, self.lstm, need_all=True)
                b, t, h = out.shape
                return self.fc(out.reshape(b * t, h)).reshape(b, t, -1).transpose(1, 2)
            return self.fc(fused.lstm2_embed_forward(input_seq, self.embeddings, self.lstm))
        out, _ = self.lstm(self.embeddings(input_seq))
        if self.per_position:
            b, t, h = out.shape
            return self.fc(out.reshape(b * t, h)).reshape(b, t, -1).transpose(1, 2)
        return self.fc(out[:, -1])


class RNN_StackOverFlow(nn.Module):
    def __init__(self, vocab_size: int = 10000, num_oov_buckets: int = 1, embedding_size: int = 96,
                 latent_size: int = 670, num_layers: int = 1):
        super().__init__()
        extended = vocab_size + 3 + num_oov_buckets  # pad/bos/eos + oov buckets
        self.word_embeddings = nn.Embedding(num_embeddings=extended, embedding_dim=embedding_size, padding_idx=0)
        self.lstm = nn.LSTM(input_size=embedding_size, hidden_size=latent_size, num_layers=num_layers)
        self.fc1 = TcLinear(latent_size, embedding_size)
        self.fc2 = TcLinear(embedding_size, extended)

    def forward(self, input_seq, hidden_state=None):
        out, hidden_state = self.lstm(self.word_embeddings(input_seq), hidden_state)
        return self.fc2(self.fc1(out[:, -1]))
