"""Label vectors `y` (reference: jukebox/data/labels.py).

y = [total_length, offset, sample_length, artist_id, genre_ids..., lyric_tokens...].
The artist / genre name tables and the lyric character vocabulary are data, not hot path: ids can
be given directly (get_y_from_ids); name lookup loads the reference's id tables from a directory
given by JUKEBOX_IDS_DIR (v2_/v3_{artist,genre}_ids.txt) and falls back to id 0 ("unknown")
exactly like the reference does for unknown names."""
import os
import re

import numpy as np
import torch as t


def get_relevant_lyric_tokens(full_tokens, n_tokens, total_length, offset, duration):
    """linear-window heuristic for which lyric characters accompany this audio window"""
    if len(full_tokens) < n_tokens:
        pad = n_tokens - len(full_tokens)
        tokens = [0] * pad + full_tokens
        indices = [-1] * pad + list(range(len(full_tokens)))
    else:
        assert 0 <= offset < total_length
        midpoint = int(len(full_tokens) * (offset + duration / 2.0) / total_length)
        midpoint = min(max(midpoint, n_tokens // 2), len(full_tokens) - n_tokens // 2)
        lo, hi = midpoint - n_tokens // 2, midpoint + n_tokens // 2
        tokens, indices = full_tokens[lo:hi], list(range(lo, hi))
    assert len(tokens) == n_tokens and len(indices) == n_tokens
    return tokens, indices


class EmptyLabeller:
    def get_label(self, artist=None, genre=None, lyrics=None, total_length=None, offset=None):
        return dict(y=np.array([], dtype=np.int64), info=dict(artist="n/a", genre="n/a", lyrics=[], full_tokens=[]))

    def get_batch_labels(self, metas, device='cpu'):
        labels = [self.get_label() for _ in metas]
        ys = t.stack([t.from_numpy(l['y']) for l in labels], dim=0).to(device).long()
        return dict(y=ys, info=[l['info'] for l in labels])


class _Vocab:
    def __init__(self, v3):
        chars = 'ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789.,:;!?-' + ('' if v3 else '+') + '\'\"()[] \t\n'
        self.vocab = {c: i + 1 for i, c in enumerate(chars)}
        self.n_vocab = len(chars) + 1
        self.tokens = {i + 1: c for i, c in enumerate(chars)}
        self.tokens[0] = ''
        self._drop = re.compile('[^' + re.escape(chars) + ']+')

    def clean(self, text):
        try:
            from unidecode import unidecode
            text = unidecode(text)
        except ImportError:
            text = text.encode('ascii', 'ignore').decode()
        return self._drop.sub('', text.replace('\\', '\n'))

    def tokenise(self, text):
        return [self.vocab[c] for c in text]

    def textise(self, tokens):
        return ''.join(self.tokens[tok] for tok in tokens)


class _NameTables:
    def __init__(self, v3):
        self.v3 = v3
        self.artist_ids, self.genre_ids = {}, {}
        d = os.environ.get("JUKEBOX_IDS_DIR")
        if d:
            ver = "v3" if v3 else "v2"
            for table, name in ((self.artist_ids, "artist"), (self.genre_ids, "genre")):
                path = os.path.join(d, f"{ver}_{name}_ids.txt")
                if os.path.exists(path):
                    with open(path, encoding="utf-8") as f:
                        for line in f:
                            key, idx = line.strip().split(';')
                            table[key.lower()] = int(idx)

    @staticmethod
    def _norm(s):
        s = ''.join(c if c.isascii() and c.isalnum() else '_' for c in s.lower())
        return re.sub(r'_+', '_', s).strip('_')

    def _lookup(self, table, key, kind):
        """id of `key`; unknown names map to id 0 ("unknown") WITH a warning, as the reference prints one
        (data/artist_genre_processor.py) - a silent 0 would give plausible but mis-conditioned samples"""
        if key in table:
            return table[key]
        if key not in ("unknown", ""):
            why = "no id tables loaded: set JUKEBOX_IDS_DIR" if not table else "not in the id table"
            print(f"Input {kind} {key!r} maps to unknown ({why})")
        return 0

    def artist(self, name):
        return self._lookup(self.artist_ids, name.lower() if self.v3 else self._norm(name), "artist")

    def genres(self, name):
        words = [name.lower()] if self.v3 else self._norm(name).split('_')
        return [self._lookup(self.genre_ids, w, "genre") for w in words]


class Labeller:
    def __init__(self, max_genre_words, n_tokens, sample_length, v3=False):
        self.names = _NameTables(v3)
        self.text = _Vocab(v3)
        self.n_tokens, self.max_genre_words, self.sample_length = n_tokens, max_genre_words, sample_length
        self.label_shape = (4 + max_genre_words + n_tokens,)

    def get_label(self, artist, genre, lyrics, total_length, offset):
        lyrics = self.text.clean(lyrics)
        full_tokens = self.text.tokenise(lyrics)
        tokens, _ = get_relevant_lyric_tokens(full_tokens, self.n_tokens, total_length, offset, self.sample_length)
        y = self.get_y_from_ids(self.names.artist(artist), self.names.genres(genre), tokens, total_length, offset)
        return dict(y=y, info=dict(artist=artist, genre=genre, lyrics=lyrics, full_tokens=full_tokens))

    def get_y_from_ids(self, artist_id, genre_ids, lyric_tokens, total_length, offset):
        assert len(genre_ids) <= self.max_genre_words
        genre_ids = list(genre_ids) + [-1] * (self.max_genre_words - len(genre_ids))
        if self.n_tokens > 0:
            assert len(lyric_tokens) == self.n_tokens
        else:
            lyric_tokens = []
        y = np.array([total_length, offset, self.sample_length, artist_id, *genre_ids, *lyric_tokens], dtype=np.int64)
        assert y.shape == self.label_shape, f"Expected {self.label_shape}, got {y.shape}"
        return y

    def get_batch_labels(self, metas, device='cpu'):
        labels = [self.get_label(**meta) for meta in metas]
        ys = t.stack([t.from_numpy(l['y']) for l in labels], dim=0).to(device).long()
        return dict(y=ys, info=[l['info'] for l in labels])

    def set_y_lyric_tokens(self, ys, labels):
        info = labels['info']
        assert ys.shape[0] == len(info)
        if self.n_tokens == 0:
            return None
        tokens_list, indices_list = [], []
        ys_host = ys.cpu()
        for i in range(ys.shape[0]):
            total_length, offset, duration = (int(v) for v in ys_host[i, :3])
            tokens, indices = get_relevant_lyric_tokens(info[i]['full_tokens'], self.n_tokens, total_length, offset, duration)
            tokens_list.append(tokens)
            indices_list.append(indices)
        ys[:, -self.n_tokens:] = t.tensor(tokens_list, dtype=t.long, device=ys.device)
        return indices_list
