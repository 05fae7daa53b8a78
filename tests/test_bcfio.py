"""include/b200_bcfio.h: the BCF2.2 binary container without htslib (BGZF blocks, header with IDX, record framing, typed
values).  No BCF written by htslib exists in the reference tree, so this suite checks what can be checked without one:
text -> BCF -> text identity on every VCF of the reference's `call` tests, a valid gzip container (Python's gzip module
inflates it), and the exact bytes of a small record derived from the specification's typed-value rules."""
import gzip
import struct

from bcftools_b200 import vcfcall
from tests import vcf_cases


def test_text_bcf_text_round_trip_of_every_reference_vcf():
    files = vcf_cases.bundle()["files"]
    n = 0
    for name, text in sorted(files.items()):
        if not (name.endswith(".vcf") or name.endswith(".out")):
            continue
        data = text.encode("latin-1")
        for level in (6, 0):                # -Ob and -Ou
            bcf = vcfcall.vcf_to_bcf(data, level)
            assert gzip.decompress(bcf)[:5] == b"BCF\x02\x02", name
            assert bcf[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"), name      # the BGZF EOF block
            assert vcfcall.bcf_to_vcf(bcf) == data, name
        n += 1
    assert n >= 40


def test_record_bytes_follow_the_specification():
    """hts-specs VCFv4.2 §6.3: typed values are one descriptor byte (length << 4 | type) plus data; int8 = 1, float = 5,
    char = 7; a missing ID is an empty string; PASS is dictionary entry 0; a flag has no value."""
    text = (b"##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n##contig=<ID=c1>\n##contig=<ID=c2>\n"
            b"##INFO=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n##INFO=<ID=AF,Number=A,Type=Float,Description=\"f\">\n"
            b"##INFO=<ID=DB,Number=0,Type=Flag,Description=\"b\">\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"g\">\n"
            b"##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"p\">\n"
            b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\ts2\n"
            b"c2\t5\t.\tA\tC\t.\tPASS\tDP=300;AF=1;DB\tGT:PL\t0/1:10,0,300\t.:.\n")
    raw = gzip.decompress(vcfcall.vcf_to_bcf(text))
    l_text = struct.unpack("<I", raw[5:9])[0]
    hdr = raw[9:9 + l_text]
    assert hdr.endswith(b"\n\x00") and b"##INFO=<ID=DP,Number=1,Type=Integer,Description=\"d\",IDX=1>" in hdr and b"##contig=<ID=c2,IDX=1>" in hdr
    rec = raw[9 + l_text:]
    l_shared, l_indiv = struct.unpack("<II", rec[:8])
    shared, indiv = rec[8:8 + l_shared], rec[8 + l_shared:8 + l_shared + l_indiv]
    assert len(rec) == 8 + l_shared + l_indiv
    exp = struct.pack("<iiiIII", 1, 4, 1, 0x7F800001, 2 << 16 | 3, 2 << 24 | 2)        # CHROM=c2, POS0=4, rlen=1, QUAL missing, 2 alleles / 3 INFO, 2 FORMAT / 2 samples
    exp += b"\x07"                                   # ID: empty string
    exp += b"\x17A\x17C"                             # alleles
    exp += b"\x11\x00"                               # FILTER: [PASS]
    exp += b"\x11\x01" + b"\x12" + struct.pack("<h", 300)       # DP (dictionary 1) = one int16
    exp += b"\x11\x02" + b"\x15" + struct.pack("<f", 1.0)       # AF (2) = one float
    exp += b"\x11\x03" + b"\x00"                     # DB (3): flag, no value
    assert shared == exp
    # FORMAT: GT (dictionary 4) two int8 per sample: 0/1 -> 2,4; "." -> missing allele then end-of-vector
    # PL (5): three int16 per sample (300 needs 16 bits): "." -> missing, then end-of-vector padding
    exp_i = b"\x11\x04" + b"\x21" + bytes([2, 4, 0, 0x81])
    exp_i += b"\x11\x05" + b"\x32" + struct.pack("<hhhHHH", 10, 0, 300, 0x8000, 0x8001, 0x8001)
    assert indiv == exp_i
    assert vcfcall.bcf_to_vcf(vcfcall.vcf_to_bcf(text)) == text


def test_long_vectors_use_the_overflow_length():
    text = (b"##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n##contig=<ID=1>\n"
            b"##INFO=<ID=X,Number=.,Type=Integer,Description=\"x\">\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
            b"1\t1\trs1\tACGTACGTACGTACGTA\t.\t12.5\t.\tX=" + b",".join(str(i).encode() for i in range(20)) + b"\n")
    raw = gzip.decompress(vcfcall.vcf_to_bcf(text))
    assert b"\xf7\x11\x11ACGTACGTACGTACGTA" in raw          # 17 characters: length 15 escapes to a typed int8 17
    assert b"\xf1\x11\x14" + bytes(range(20)) in raw        # 20 int8 values
    assert vcfcall.bcf_to_vcf(vcfcall.vcf_to_bcf(text)) == text
