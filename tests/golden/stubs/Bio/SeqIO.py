"""GOLDEN-GENERATION STUB: the two Biopython features the reference uses (SeqIO.parse of FASTA, Seq.transcribe)."""


class Seq(str):
    def transcribe(self):
        return Seq(str(self).replace("T", "U").replace("t", "u"))

    def __getitem__(self, k):
        return Seq(str.__getitem__(self, k))


class SeqRecord:
    def __init__(self, name, seq):
        self.name = self.id = name
        self.description = name
        self.seq = Seq(seq)

    def __len__(self):
        return len(self.seq)


def parse(handle, fmt):
    close = False
    if isinstance(handle, str):
        handle = open(handle)
        close = True
    name, chunks = None, []
    for line in handle:
        line = line.rstrip("\n")
        if line.startswith(">"):
            if name is not None:
                yield SeqRecord(name, "".join(chunks))
            name = line[1:].split()[0] if line[1:].split() else ""
            chunks = []
        elif name is not None:
            chunks.append(line.strip())
    if name is not None:
        yield SeqRecord(name, "".join(chunks))
    if close:
        handle.close()
