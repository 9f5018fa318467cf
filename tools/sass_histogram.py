"""Static SASS opcode histogram of one kernel in an object file (runs without a GPU).

    python tools/sass_histogram.py <object or .so> <substring of the mangled kernel name> [--top N]

Used to check, before spending GPU time, that a change has the intended effect on the instruction stream
(e.g. that FADD2 / FFMA2 / UTMALDG appear), and to produce the histograms under profiles/."""
import collections
import re
import subprocess
import sys


def histogram(path, needle):
    text = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, universal_newlines=True).stdout
    counts, name = {}, None
    for line in text.splitlines():
        match = re.match(r"\s+Function : (\S+)", line)
        if match:
            name = match.group(1)
            continue
        match = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if match and name and needle in name:
            counts.setdefault(name, collections.Counter())[match.group(1).split(".")[0]] += 1
    return counts


def main():
    path, needle = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    for name, counter in histogram(path, needle).items():
        print(name, "total", sum(counter.values()))
        for op, n in counter.most_common(top):
            print("  %-12s %d" % (op, n))


if __name__ == "__main__":
    main()
