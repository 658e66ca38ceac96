"""Known-answer vectors restated (as data) from the reference's own tests and docs.

Each entry cites where the reference pins it (paths relative to /root/reference/).  The same
vectors are replayed against the CPU oracle (-m "not gpu") and against the CUDA path (-m gpu).
"""

TABLES_DAT_SHA256 = "aedba937837335b5999364aa43b47b7f32fb381e6162aca9f6f4209b3ca687bd"  # src/tables.dat

# tests/utils.py:7-12 -- tables the reference's conformance tests run
VALID_TABLES = [t for t in range(1, 34) if t not in (7, 8, 17, 18, 19, 20)]
ALL_ACCEPTED_TABLES = [t for t in range(1, 34) if t not in (17, 18, 19, 20)]  # src/trans_table.rs:230-267

# (table, dna, protein)
TRANSLATE = [
    (1, "AAAGGGAAA", "KGK"),  # src/rust_api.rs:460-469, tests/test_wrapper.py:9-10
    (1, "TTRTTV", "LX"),  # src/rust_api.rs:472-482 ("TTR TTV" after the parse drops the space)
    (1, "GGNATN", "GX"),  # README.md:31-37
    (1, "RAT", "B"),  # README.md:31-37
    (1, "taatcaagactattcaaccaa", "*SRLFNQ"),  # README.md:18
    (22, "taatcaagactattcaaccaa", "**RLFNQ"),  # README.md:64-71
    (1, "ATCGATCG", "ID"),  # src/trans_table.rs:155-164
    (1, "CCNTACACKCATNCNAAT", "PYTHXN"),  # src/python_api.rs:32 (doc, without the stale space)
    (1, "AAACCCTTTGGG", "KPFG"),  # src/python_api.rs:44
    (1, "GGG", "G"),  # src/rust_api.rs:667-679
    (1, "AAAQ", "K"),  # trailing n%3 bytes are never validated (src/trans_table.rs:191)
    (1, "", ""),  # src/trans_table.rs:180-182
    (1, "AA", ""),
]

# (dna, [frames]) -- table 1
SELF_FRAMES = [
    ("AAAGGGAAA", ["KGK", "KG", "RE"]),  # src/rust_api.rs:484-496, tests/test_wrapper.py:13-18
    ("GGGG", ["G", "G"]),  # src/rust_api.rs:498-526
    ("GGG", ["G"]),
    ("GG", []),
    ("G", []),
    ("", []),
]
ALL_FRAMES = [
    ("AAAGGGAAA", ["KGK", "KG", "RE", "FPF", "FP", "SL"]),  # src/rust_api.rs:556-578, tests/test_wrapper.py:35-44
    ("GGGG", ["G", "G", "P", "P"]),  # src/rust_api.rs:580-632, tests/test_wrapper.py:47-62
    ("GGG", ["G", "P"]),
    ("GG", []),
    ("G", []),
    ("", []),
    ("taatcaagactattcaaccaa", ["*SRLFNQ", "NQDYST", "IKTIQP", "LVE*S*L", "WLNSLD", "G*IVLI"]),  # README.md:46-52
    ("AAAA", ["K", "K", "F", "F"]),  # README.md:53-55 (4 frames)
    ("AA", []),
]

# codon-level frame structure of the lazy API, src/iter.rs:21-74 (dna = CGATCGAT)
ITER_FRAMES_DNA = "CGATCGAT"
ITER_FRAMES_CODONS = [["CGA", "TCG"], ["GAT", "CGA"], ["ATC", "GAT"], ["ATC", "GAT"], ["TCG", "ATC"], ["CGA", "TCG"]]
ITER_FRAMES_DNA4_CODONS = [["CGA"], ["GAT"], ["ATC"], ["TCG"]]  # dna[..4]

REVCOMP = [
    ("tcaaga", "TCTTGA"),  # README.md:21-22
    ("AAAGGGGGG", "CCCCCCTTT"),  # tests/test_wrapper.py:143-145
    ("AAABBBGGG", "CCCVVVTTT"),  # tests/test_wrapper.py:146-147
    ("AAAAABCCC", "GGGVTTTTT"),  # src/python_api.rs:57
    ("AAAAAACCC", "GGGTTTTTT"),  # src/python_api.rs:68
    ("CGAT", "ATCG"),  # src/iter.rs:163-170
    ("", ""),
]
RC_FRAME0 = [("CATTAG", "LM")]  # tests/integration_tests.rs:24-39

# (callable name, args, kind, payload byte, message)
ERRORS = [
    ("revcomp_strict", "AAABBBGGG", 3, ord("B"), "unexpected ambiguous nucleotide: 'B'"),  # tests/test_wrapper.py:148-149
    ("revcomp_strict", "AAAAAACCN", 3, ord("N"), "unexpected ambiguous nucleotide: 'N'"),  # src/python_api.rs:69
    ("translate_strict", "AAABBBAAA", 3, ord("B"), "unexpected ambiguous nucleotide: 'B'"),  # tests/test_wrapper.py:152-157
    ("translate_strict", "AAACCCTTTGGN", 3, ord("N"), "unexpected ambiguous nucleotide: 'N'"),  # src/python_api.rs:45
    ("translate_strict", "ABCD", 3, ord("B"), "unexpected ambiguous nucleotide: 'B'"),  # src/fasta.rs:1156-1169
    ("translate", "AAAelephant", 2, ord("e"), "bad nucleotide: 'e'"),  # src/fasta.rs:1230-1260
    ("translate", "AA\xc4\x8dCCG", 1, 0xC4, "non-ascii byte: c4"),  # src/fasta.rs:1262-1288 (AAčCCG, utf-8)
    ("translate", "AAxAAA", 2, ord("x"), "bad nucleotide: 'x'"),  # src/fasta.rs:1388-1396
    ("translate", "CCNTACACK CATNCNAAT", 2, ord(" "), "bad nucleotide: ' '"),  # bytes path skips no whitespace
    ("translate_strict", "AAb", 3, ord("B"), "unexpected ambiguous nucleotide: 'B'"),  # upper-cased payload
]
BAD_TABLE_MESSAGE = "not a ncbi translation table: 17"  # src/errors.rs:27-28

# src/rust_api.rs:421-457
ACCEPTED_AMBIGUOUS = "aAtTcCgGmMrRwWsSyYkKvVhHdDbBnN"
ACCEPTED_STRICT = "aAtTcCgG"

CANONICAL = [
    ("CATTAG", "ATCCTG"),  # src/rust_api.rs:365-367
    ("TGCGAGTGTAGCGAGATGTAGCGTAGAGTCTGAGATGCAGTA", "ATCAGCTACACTGTCACATCGCATCTACACGCATCTCACGCT"),  # src/canonical.rs:208-214
    ("TTGT", "AATA"),  # src/canonical.rs:406-412
    ("TGTT", "AATA"),
    ("ATCGCCAT", "ATCCGCAT"),
    ("", ""),
]
FORWARD_CANONICAL = [
    ("TGCGAGTGTAGCGAGATGTAGCGTAGAGTCTGAGATGCAGTA", "ATCTGTATAGTCTGTGATAGTCTAGTGTACATGTGATCGTAG"),  # src/canonical.rs:200-206
    ("CATTAG", "ATCCTG"),  # src/canonical.rs:66-67 (doc)
]
# src/canonical.rs:216-404: all 85 inputs of length <= 3; value = (forward_canonical, canonical).
# Encoded compactly: for inputs over ATCG in the reference's listing order (A,T,C,G per position).
SMALL_FW = {
    1: "A A A A",
    2: "AA AT AT AT AT AA AT AT AT AT AA AT AT AT AT AA",
    3: ("AAA AAT AAT AAT ATA ATT ATC ATC ATA ATC ATT ATC ATA ATC ATC ATT "
        "ATT ATA ATC ATC AAT AAA AAT AAT ATC ATA ATT ATC ATC ATA ATC ATT "
        "ATT ATC ATA ATC ATC ATT ATA ATC AAT AAT AAA AAT ATC ATC ATA ATT "
        "ATT ATC ATC ATA ATC ATT ATC ATA ATC ATC ATT ATA AAT AAT AAT AAA"),
}
SMALL_CANON = {
    1: "A A A A",
    2: "AA AT AT AT AT AA AT AT AT AT AA AT AT AT AT AA",
    3: ("AAA AAT AAT AAT ATA AAT ATC ATC ATA ATC AAT ATC ATA ATC ATC AAT "
        "AAT ATA ATC ATC AAT AAA AAT AAT ATC ATA AAT ATC ATC ATA ATC AAT "
        "AAT ATC ATA ATC ATC AAT ATA ATC AAT AAT AAA AAT ATC ATC ATA AAT "
        "AAT ATC ATC ATA ATC AAT ATC ATA ATC ATC AAT ATA AAT AAT AAT AAA"),
}

EXPANSIONS = [
    ("ATBCGYAC", ["ATTCGTAC", "ATTCGCAC", "ATCCGTAC", "ATCCGCAC", "ATGCGTAC", "ATGCGCAC"]),  # src/expansions.rs:208-226
    ("WWW", ["AAA", "AAT", "ATA", "ATT", "TAA", "TAT", "TTA", "TTT"]),  # :228-233
    ("ATCGNGCTA", ["ATCGAGCTA", "ATCGTGCTA", "ATCGCGCTA", "ATCGGGCTA"]),  # :294-305
    ("ATCGATATCGCGAATTCCGG", ["ATCGATATCGCGAATTCCGG"]),  # :273-281
]
# src/expansions.rs:235-271
SIZE_HINTS = [("", 1), ("A", 1), ("ATCG", 1), ("W", 2), ("B", 3), ("N", 4), ("MR", 4), ("YV", 6), ("DS", 6),
              ("KN", 8), ("NW", 8), ("MRY", 8), ("HB", 9), ("ACGTYTAGCNCGATVGCTA", 24)]
# src/expansions.rs:283-292
SIZE_HINT_BIG = ("NNNNNNNNNNNNNNNNNNBBBBBBBBBBBBBBBBBY", 17748888853923495936)
SIZE_HINT_OVERFLOW = "NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN"

# src/rust_api.rs:702-769
WINDOWS_10 = ("gcantacctaangtnattag",
              ["gcantaccta", "cantacctaa", "antacctaan", "ntacctaang", "tacctaangt", "acctaangtn", "cctaangtna",
               "ctaangtnat", "taangtnatt", "aangtnatta", "angtnattag"])

# Derived pins recorded by the survey with an independent model for
# benchmarks/covid.txt with newlines stripped (29 901 nt).  Checked only where the file exists.
COVID = {
    "len": 29901,
    "sha256_prefix": "ab3e4f3e900dedc9",
    "translate_t1_sha256": "e4c87be3e601ee04156595fc5a4cada4c9ed35af8047f4805c1e4242d889baf3",
    "translate_t1_head": "IKGLYLPR*QTNQLSISCRSVL*TNFKICVAVTRLHA*CT",
    "translate_t2_sha256_prefix": "97914611f0bd8a79",
    "translate_t22_sha256_prefix": "5139e5b63974c535",
    "revcomp_sha256": "cb628e8bbfb51a5cfb564157d98e9127ef1160ebfd362c3ee0a7ad2f1a441255",
    "frame_lens": [9967, 9966, 9966, 9967, 9966, 9966],
    "frames_joined_sha256": "c050477ed99d645826c4e073904b843ee5fbf723a38dc7bc680c0518f4d75545",
}


# ---- FASTA (src/fasta.rs tests, restated with DNA content): (text, settings or None for all four, expected records)
# a record = (header, content letters, (line_begin, line_end)); settings = (concatenate_headers, allow_preceding_comment)
_ALL = [(c, a) for c in (False, True) for a in (False, True)]
FASTA = [
    ("", _ALL, []),                                                                  # :598-601
    ("\n\n", _ALL, []),                                                              # :603-606
    ("  \n\n \n  \t\t \t\n  \t \n", _ALL, []),                                       # :608-611 with the DNA-legal blanks
    (">Virus\n", _ALL, [("Virus", "", (1, 2))]),                                     # :613-633
    (";Virus\n", _ALL, [("Virus", "", (1, 2))]),
    (">Virus\n\n", _ALL, [("Virus", "", (1, 3))]),                                   # :635-654
    ("ACGT\nGGCC", [(False, True), (True, True)], []),                               # :656-663 preceding content dropped
    ("ACGT\nGGCC", [(False, False), (True, False)], [("", "ACGTGGCC", (1, 3))]),     # :665-676
    ("ACGT\nGGCC\n\n>Virus\n\n", [(False, True), (True, True)], [("Virus", "", (4, 6))]),  # :678-690
    ("ACGT\nGGCC\n\n>Virus\n\n", [(False, False), (True, False)], [("", "ACGTGGCC", (1, 4)), ("Virus", "", (4, 6))]),  # :692-712
    ("   \t\n\t\t   \n\n>Virus\n\n", _ALL, [("Virus", "", (4, 6))]),                # :714-724 (blank preceding content)
    (">Virus\nCAAAGT\n", _ALL, [("Virus", "CAAAGT", (1, 3))]),                       # :726-746
    (";Virus\nCAAAGT", _ALL, [("Virus", "CAAAGT", (1, 3))]),                         # :748-768 (no trailing newline)
    (">Virus\nAAAA\nCCCC\nGGGG\n", _ALL, [("Virus", "AAAACCCCGGGG", (1, 5))]),       # :770-790
    (">Virus1\nAAAA\n;Virus2\nCCCC\n", _ALL, [("Virus1", "AAAA", (1, 3)), ("Virus2", "CCCC", (3, 5))]),  # :792-853
    (">Virus1\nAAAA\nAAAA\n>Virus2\nCCCC\nCCCC\n>Virus3\nCCCC\nCCCC\n", _ALL,
     [("Virus1", "AAAAAAAA", (1, 4)), ("Virus2", "CCCCCCCC", (4, 7)), ("Virus3", "CCCCCCCC", (7, 10))]),  # :920-943
    (">Virus1\nAAAA\nAAAA\n>Virus2\nCCCC\nCCCC\n>Virus3", _ALL,
     [("Virus1", "AAAAAAAA", (1, 4)), ("Virus2", "CCCCCCCC", (4, 7)), ("Virus3", "", (7, 8))]),           # :945-968
    (">a\n>b\ntcga", [(True, False), (True, True)], [("a\nb", "TCGA", (1, 4))]),     # :970-981
    (">a\n>b\ntcga", [(False, False), (False, True)], [("a", "", (1, 2)), ("b", "TCGA", (2, 4))]),  # :983-1003
    (">a\n>b", [(True, False), (True, True)], [("a\nb", "", (1, 3))]),               # :1005-1016
    (">Virus1\r\nAC GT\r\n\r\n>V2 x\r\nTT\tT\r\n", _ALL, [("Virus1", "ACGT", (1, 4)), ("V2 x", "TTT", (4, 6))]),  # CRLF (BufRead::lines)
    (">Virus1\nAAAA \n", _ALL, [("Virus1", "AAAA", (1, 3))]),                        # :1213-1222
    (">Virus1\n  AAAA\tBCD \t\n", _ALL, [("Virus1", "AAAABCD", (1, 3))]),            # :1223-1231
    (">Virus1\nAC\nT\n>Empty\n\n>Virus2\n>with many\n>comment lines\nC  AT", [(True, False)],
     [("Virus1", "ACT", (1, 4)), ("Empty", "", (4, 6)), ("Virus2\nwith many\ncomment lines", "CAT", (6, 10))]),  # :1342-1358
]
# (text, strict, line, kind, byte, located message)
FASTA_ERRORS = [
    (">Virus1\nABCD", True, 2, 3, ord("B"), "on line 2: error parsing record: unexpected ambiguous nucleotide: 'B'"),   # :1157-1169
    (">Virus1\nAAAelephant", False, 2, 2, ord("e"), "on line 2: error parsing record: bad nucleotide: 'e'"),           # :1233-1250
    (">Virus1\nAAAA\n>Virus2\nAAAAelephant", True, 4, 2, ord("e"), "on line 4: error parsing record: bad nucleotide: 'e'"),  # :1252-1270
    (">Virus1\nAAčCCG\n", False, 2, 1, 196, "on line 2: error parsing record: non-ascii byte: c4"),                # :1272-1290
    (">Virus1\nAAA\nCCCxGGG", True, 3, 2, ord("x"), "on line 3: error parsing record: bad nucleotide: 'x'"),           # :1388-1396
    ("AC\rGT\n>x\n", False, 1, 2, 13, "on line 1: error parsing record: bad nucleotide: '\\r'"),                        # a bare CR is content
]
