"""Stand-in for ``pyahocorasick`` (>= 2.1, absent here).  TEST INFRASTRUCTURE ONLY.

Provides what the reference calls (``bin/getCandidateKmers.py:112,152,155,171,
315,406-409``): ``Automaton.add_word / make_automaton / iter``.

Recalled semantics of ``Automaton.iter(s)``: yields ``(end_index, value)`` for
every occurrence of every stored word, by increasing end index; at one end
index the longest word comes first (the search walks fail links from the
deepest matching trie node).  Not verifiable against a real install here.
"""


class Automaton:
    def __init__(self):
        self._by_len = {}
        self._lens = []

    def add_word(self, word, value):
        self._by_len.setdefault(len(word), {})[word] = value
        return True

    def make_automaton(self):
        self._lens = sorted(self._by_len.keys(), reverse=True)

    def __len__(self):
        return sum(len(d) for d in self._by_len.values())

    def iter(self, s):
        if not self._lens and self._by_len:
            self.make_automaton()
        lens = [(n, self._by_len[n]) for n in self._lens]
        for end in range(len(s)):
            for n, words in lens:
                st = end - n + 1
                if st < 0:
                    continue
                v = words.get(s[st:end + 1])
                if v is not None:
                    yield end, v

    # pickling support (the reference ships the homopolymer automaton to Pool workers)
    def __getstate__(self):
        return {"by_len": self._by_len, "lens": self._lens}

    def __setstate__(self, st):
        self._by_len = st["by_len"]
        self._lens = st["lens"]
