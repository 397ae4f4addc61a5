"""TEST INFRASTRUCTURE ONLY. Byte-compiles the reference's own test file from where it lies into
oracle/_ref/ (usage: compile_reference_tests.py <source.py> <out.marshal>). The output is a header line
(python version + sha256 of the source) followed by the marshalled code object; nothing of the source
text is stored. tests/test_gpu_reference_file.py loads and runs it unchanged."""
import hashlib
import marshal
import sys


def main(src: str, out: str) -> None:
    text = open(src, "rb").read()
    code = compile(text, "reference/tests/test_array.py", "exec", dont_inherit=True)
    header = f"{sys.version_info[0]}.{sys.version_info[1]} {hashlib.sha256(text).hexdigest()}\n".encode()
    with open(out, "wb") as f:
        f.write(header + marshal.dumps(code))
    print(f"compiled {src} -> {out}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
