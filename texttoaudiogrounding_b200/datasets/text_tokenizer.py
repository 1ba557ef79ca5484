"""DictTokenizer — host-side mirror of reference datasets/text_tokenizer.py:9-58 (+ the Vocabulary wrapper of
utils/build_vocab.py:7-34): whitespace tokens -> vocabulary ids, unknown words -> ``<unk>``, zero padding.
Accepts List[str] ([B, N] ids) or List[List[str]] ([B, text_num, N] ids, the multi-phrase schema)."""
from __future__ import annotations

import pickle
from typing import Dict, List, Union

import numpy as np
import torch


class Vocabulary:
    def __init__(self):
        self.word2idx = {}
        self.idx2word = {}
        self.idx = 0

    def add_word(self, word):
        if word not in self.word2idx:
            self.word2idx[word] = self.idx
            self.idx2word[self.idx] = word
            self.idx += 1

    def __call__(self, word):
        if word not in self.word2idx:
            return self.word2idx["<unk>"]
        return self.word2idx[word]

    def __len__(self):
        return len(self.word2idx)

    def state_dict(self):
        return self.word2idx

    def load_state_dict(self, state_dict):
        self.word2idx = state_dict
        self.idx2word = {idx: word for word, idx in self.word2idx.items()}
        self.idx = len(self.word2idx)


def pad_sequence(data):
    """utils/train_util.py:211-216"""
    data = [torch.as_tensor(arr) for arr in data]
    padded = torch.nn.utils.rnn.pad_sequence(data, batch_first=True)
    length = torch.as_tensor([x.shape[0] for x in data]).long()
    return padded, length


class DictTokenizer:
    def __init__(self, vocabulary: Union[str, Dict[str, int]]) -> None:
        self.vocabulary = Vocabulary()
        state_dict = pickle.load(open(vocabulary, "rb")) if isinstance(vocabulary, str) else dict(vocabulary)
        self.vocabulary.load_state_dict(state_dict)

    def _encode(self, texts: List[str]):
        return pad_sequence([np.array([self.vocabulary(tok) for tok in t.split()], dtype=np.int64) for t in texts])

    def __call__(self, texts):
        assert isinstance(texts, list), "the input must be List[str] or List[List[str]]"
        if isinstance(texts[0], str):
            tokens, token_lens = self._encode(texts)
        else:
            text_num, batch_size = len(texts[0]), len(texts)
            for text_list in texts:
                assert len(text_list) == text_num, "the text number in each list must be the same"
            tokens, token_lens = self._encode(sum(texts, []))
            tokens = tokens.reshape(batch_size, text_num, -1)
            token_lens = token_lens.reshape(batch_size, text_num)
        return {"text": tokens, "text_len": token_lens}

    def inverse_transform(self, texts):
        output = []
        for text in texts:
            tokens = []
            for word_idx in text:
                if word_idx != 0:
                    tokens.append(self.vocabulary.idx2word[int(word_idx)])
                else:
                    break
            output.append(" ".join(tokens))
        return output
