/*  b200_vcfcall.c -- the `bcftools call -m` driver around the B200 path; see include/b200_vcfcall.h.
 *  Reference line map: option parsing vcfcall.c:945-1062, init_data 608-712, next_line 471-606, main loop 1089-1156,
 *  tgt_* 346-455, set_ploidy 807-825; record edits of mcall() mcall.c:1430-1460, 1536-1543, 1576-1684,
 *  mcall_trim_and_update_numberR 1196-1265, mcall_constrain_alleles 1271-1421; gvcf.c:88-227; vcmp.c:55-131.  */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <stdarg.h>
#include <ctype.h>
#include "b200_vcfcall.h"
#include "b200_driver.h"
#include "b200_bcfio.h"

#define CF_INS_MISSED   (1<<4)      /* vcfcall.c:62-69 */
#define CF_INDEL_ONLY   1
#define CF_NO_INDEL     (1<<1)
#define CF_ACGT_ONLY    (1<<2)
#define TGT_MAX         0xffffffffu  /* REGIDX_MAX */

typedef struct { char *chr; uint32_t pos; int n; char **allele; int used; } tgt_t;     /* one line of the -T file (tgt_als_t + position) */

typedef struct
{
    int *dp_range, ndp_range, prev_range;
    int32_t *dp, *pl, *tmp, *gts; int mdp, mpl, npl, mtmp, ngts, mgts;
    char *chr; int64_t start, end; int32_t min_dp;
    b200_str_t als;
}
gvcf_t;

struct b200_vcrec
{
    b200_vrec_t *rec;
    char *pre; size_t npre;         /* -i: lines written in front of this record */
    int unseen, nals_ori;
    int passthrough;                /* written without calling (too many alleles) */
    int32_t *PLs; int nPLs, mPLs;
    float *QS; int nQS, mQS;
    int32_t *ADs; int nADs, mADs;
    int32_t *prior_ac; int n_prior_ac, m_prior_ac;
    uint8_t *ploidy;                /* the record's ploidy vector (mcall_trim_and_update_PLs is per-ploidy on the device already) */
};

struct b200_vc
{
    /* options */
    uint32_t aux_flag, output_tags; int flag;
    double theta;
    char *prior_AN, *prior_AC, *grp_tag;
    int grouped;
    /* input */
    const char *text; size_t len, off;
    b200_vhdr_t *hdr;
    int nsmpl;
    /* samples, ploidy, groups */
    b200_ploidy_t *ploidy; int nsex, *sample2sex, *sex2ploidy_prev; uint8_t *ploidy_vec;
    uint32_t *grp_off, *grp_smpl; int ngroups;
    /* targets */
    tgt_t *tgt; int ntgt; char **tgt_chr; int ntgt_chr;
    int have_prev; char *prev_chr; uint32_t prev_beg;
    tgt_t *cur_tgt;                 /* aux.tgt_als of the current record */
    b200_vrec_t **buf; int nbuf, mbuf;     /* -C alleles: records at duplicate positions (vcfbuf) */
    int eof;
    gvcf_t *gvcf;
    int unseen;
    int output_type;                /* 'v' text, 'z' BGZF-compressed text, 'b' BCF, 'u' uncompressed BCF (vcfcall.c:1003-1011) */
    b200_str_t out;
    char err[512];
};

static int vc_fail(b200_vc_t *vc, const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(vc->err, sizeof vc->err, fmt, ap);
    va_end(ap);
    return -1;
}
const char *b200_vc_error(const b200_vc_t *vc) { return vc->err; }

static char *read_file(const char *path, size_t *len)
{
    FILE *fp = fopen(path, "rb");
    if ( !fp ) return NULL;
    b200_str_t s = {0,0,0};
    char buf[65536]; size_t n;
    while ( (n = fread(buf, 1, sizeof buf, fp)) > 0 ) b200_str_putsn(&s, buf, n);
    fclose(fp);
    if ( !s.s ) s.s = (char*) calloc(1, 1);
    if ( len ) *len = s.l;
    return s.s;
}
static char *xstrdup(const char *s) { size_t n = strlen(s); char *d = (char*) malloc(n+1); if ( d ) memcpy(d, s, n+1); return d; }

/* ---- vcmp (vcmp.c:55-131) --------------------------------------------------------------------------- */
int b200_vcmp_set_ref(const char *ref1, const char *ref2, char *dref, size_t mdref, int *ndref)
{
    *ndref = 0;
    const char *a = ref1, *b = ref2;
    while ( *a && *b && toupper((unsigned char)*a)==toupper((unsigned char)*b) ) { a++; b++; }
    if ( !*a && !*b ) return 0;
    if ( *a && *b ) return -1;
    const char *lng = *a ? ref1 : ref2, *rest = *a ? a : b;
    int nmatch = (int)(rest - lng), n = (int) strlen(rest);
    if ( (size_t)n + 1 > mdref ) return -1;
    for (int i=0; i<n; i++) dref[i] = (char) toupper((unsigned char)lng[nmatch+i]);
    dref[n] = 0;
    *ndref = *a ? n : -n;       /* positive when ref1 is longer */
    return 0;
}
int b200_vcmp_find_allele(const char *dref, int ndref, const char *const *als1, int nals1, const char *al2)
{
    int i, j;
    for (i=0; i<nals1; i++)
    {
        const char *a = als1[i], *b = al2;
        while ( *a && *b && toupper((unsigned char)*a)==toupper((unsigned char)*b) ) { a++; b++; }
        if ( *a && *b ) continue;
        if ( !ndref )
        {
            if ( !*a && !*b ) break;
            continue;
        }
        if ( *a )
        {
            if ( ndref<0 ) continue;
            for (j=0; j<ndref; j++)
                if ( !a[j] || toupper((unsigned char)a[j])!=dref[j] ) break;
            if ( j!=ndref || a[j] ) continue;
            break;
        }
        if ( ndref>0 ) continue;
        for (j=0; j<-ndref; j++)
            if ( !b[j] || toupper((unsigned char)b[j])!=dref[j] ) break;
        if ( j!=-ndref || b[j] ) continue;
        break;
    }
    return i==nals1 ? -1 : i;
}

/* ---- targets (vcfcall.c:359-455) -------------------------------------------------------------------- */
static int tgt_cmp(const void *a, const void *b)
{
    const tgt_t *x = (const tgt_t*)a, *y = (const tgt_t*)b;
    int c = strcmp(x->chr, y->chr);
    if ( c ) return c;
    return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}
static int tgt_load(b200_vc_t *vc, const char *path, int with_alleles)
{
    char *txt = read_file(path, NULL);
    if ( !txt ) return vc_fail(vc, "Could not read the targets file: %s\n", path);
    char *p = txt;
    int m = 0;
    while ( *p )
    {
        char *e = strchr(p, '\n');
        if ( e ) *e = 0;
        char *ss = p;
        while ( *ss && isspace((unsigned char)*ss) ) ss++;
        if ( *ss && *ss!='#' )
        {
            char *se = ss;
            while ( *se && !isspace((unsigned char)*se) ) se++;
            if ( !*se ) { free(txt); return vc_fail(vc, "Could not parse the line: %s\n", p); }
            if ( vc->ntgt==m ) { m = m ? 2*m : 64; vc->tgt = (tgt_t*) realloc(vc->tgt, sizeof(tgt_t)*m); }
            tgt_t *t = &vc->tgt[vc->ntgt];
            memset(t, 0, sizeof *t);
            *se = 0; t->chr = xstrdup(ss);
            ss = se+1;
            double beg = strtod(ss, &se);
            if ( ss==se || beg==0 ) { free(txt); return vc_fail(vc, "Could not parse tab line, expected 1-based coordinate: %s\n", p); }
            t->pos = (uint32_t)beg - 1;
            if ( with_alleles )
            {
                ss = se;
                while ( *ss && isspace((unsigned char)*ss) ) ss++;     /* the reference takes se+1: one separator */
                while ( *ss )
                {
                    se = ss;
                    while ( *se && *se!=',' && *se!='\r' ) se++;
                    char c = *se; *se = 0;
                    t->allele = (char**) realloc(t->allele, sizeof(char*)*(t->n+1));
                    t->allele[t->n++] = xstrdup(ss);
                    if ( c!=',' ) break;
                    ss = se+1;
                }
            }
            /* chromosome names in the order of their first line (regidx_seq_names) */
            int k;
            for (k=0; k<vc->ntgt_chr; k++) if ( !strcmp(vc->tgt_chr[k], t->chr) ) break;
            if ( k==vc->ntgt_chr )
            {
                vc->tgt_chr = (char**) realloc(vc->tgt_chr, sizeof(char*)*(vc->ntgt_chr+1));
                vc->tgt_chr[vc->ntgt_chr++] = xstrdup(t->chr);
            }
            vc->ntgt++;
        }
        if ( !e ) break;
        p = e+1;
    }
    free(txt);
    /* position order within a chromosome, file order among equal positions (the index of regidx.c) */
    for (int i=1; i<vc->ntgt; i++)      /* stable insertion sort: target files are short */
    {
        tgt_t t = vc->tgt[i]; int j = i;
        while ( j>0 && tgt_cmp(&vc->tgt[j-1], &t) > 0 ) { vc->tgt[j] = vc->tgt[j-1]; j--; }
        vc->tgt[j] = t;
    }
    return 0;
}
static int tgt_has_pos(const b200_vc_t *vc, const char *chr, uint32_t pos)
{
    for (int i=0; i<vc->ntgt; i++)
        if ( vc->tgt[i].pos==pos && !strcmp(vc->tgt[i].chr, chr) ) return 1;
    return 0;
}
static void missed_line(b200_vc_t *vc, b200_str_t *out, const tgt_t *t)      /* init_missed_line + tgt_flush_region, vcfcall.c:346-357, 420-428 */
{
    b200_str_puts(out, t->chr); b200_str_putc(out, '\t');
    b200_str_putw(out, (long long)t->pos + 1);
    b200_str_puts(out, "\t.\t");
    b200_str_puts(out, t->n ? t->allele[0] : "."); b200_str_putc(out, '\t');
    if ( t->n > 1 ) for (int i=1; i<t->n; i++) { if ( i>1 ) b200_str_putc(out, ','); b200_str_puts(out, t->allele[i]); }
    else b200_str_putc(out, '.');
    b200_str_puts(out, "\t.\t.\t.");
    if ( vc->nsmpl )
    {
        b200_str_puts(out, "\tGT");
        for (int i=0; i<vc->nsmpl; i++) b200_str_puts(out, "\t.");
    }
    b200_str_putc(out, '\n');
}
static void tgt_flush_region(b200_vc_t *vc, b200_str_t *out, const char *chr, uint32_t beg, uint32_t end)
{
    for (int i=0; i<vc->ntgt; i++)
    {
        tgt_t *t = &vc->tgt[i];
        if ( strcmp(t->chr, chr) || t->pos < beg || t->pos > end || t->used ) continue;
        t->used = 1;
        missed_line(vc, out, t);
    }
}
static void tgt_flush(b200_vc_t *vc, b200_str_t *out, const b200_vrec_t *rec)       /* vcfcall.c:430-455 */
{
    if ( rec )
    {
        uint32_t before = (uint32_t)rec->pos - 1u;
        if ( !vc->have_prev ) tgt_flush_region(vc, out, rec->chrom, 0, before);
        else if ( strcmp(rec->chrom, vc->prev_chr) )
        {
            tgt_flush_region(vc, out, vc->prev_chr, vc->prev_beg+1, TGT_MAX);
            tgt_flush_region(vc, out, rec->chrom, 0, before);
        }
        else tgt_flush_region(vc, out, vc->prev_chr, vc->prev_beg, before);
    }
    else
    {
        if ( vc->have_prev ) tgt_flush_region(vc, out, vc->prev_chr, vc->prev_beg, TGT_MAX);
        for (int i=0; i<vc->ntgt_chr; i++) tgt_flush_region(vc, out, vc->tgt_chr[i], 0, TGT_MAX);
    }
}
static int is_indel_als(int nals, char **als)      /* vcfcall.c:456-470 */
{
    if ( nals>1 && als[1][0]=='<' ) return 0;
    for (int i=0; i<nals; i++)
    {
        if ( als[i][0]=='<' ) continue;
        if ( als[i][1] ) return 1;
    }
    return 0;
}
static int rec_is_snp(const b200_vrec_t *v)        /* [htslib] bcf_is_snp */
{
    int i;
    for (i=0; i<v->n_allele; i++)
    {
        const char *a = v->allele[i];
        if ( a[1]==0 && a[0]!='*' ) continue;
        if ( a[0]=='<' && a[1]=='X' && a[2]=='>' ) continue;
        if ( a[0]=='<' && a[1]=='*' && a[2]=='>' ) continue;
        break;
    }
    return i==v->n_allele;
}

/* ---- options and set-up ----------------------------------------------------------------------------- */
static uint32_t parse_output_tags(const char *str)      /* vcfcall.c:764-797 */
{
    uint32_t flag = 0;
    const char *ss = str;
    while ( *ss )
    {
        const char *se = ss;
        while ( *se && *se!=',' ) se++;
        size_t n = (size_t)(se-ss);
        if ( (n==2 && !strncasecmp(ss, "GQ", 2)) || (n==6 && !strncasecmp(ss, "FMT/GQ", 6)) || (n==9 && !strncasecmp(ss, "FORMAT/GQ", 9)) ) flag |= CALL_FMT_GQ;
        else if ( (n==2 && !strncasecmp(ss, "GP", 2)) || (n==6 && !strncasecmp(ss, "FMT/GP", 6)) || (n==9 && !strncasecmp(ss, "FORMAT/GP", 9)) ) flag |= CALL_FMT_GP;
        else if ( (n==3 && !strncasecmp(ss, "PV4", 3)) || (n==8 && !strncasecmp(ss, "INFO/PV4", 8)) ) flag |= CALL_FMT_PV4;
        else return ~0u;
        if ( !*se ) break;
        ss = se+1;
    }
    return flag;
}

b200_vc_t *b200_vc_open(int argc, const char *const *argv, const char *vcf_text, size_t len, char *err, size_t errlen)
{
    b200_vc_t *vc = (b200_vc_t*) calloc(1, sizeof *vc);
    if ( !vc ) return NULL;
    #define OPEN_FAIL(...) do { vc_fail(vc, __VA_ARGS__); if ( err && errlen ) snprintf(err, errlen, "%s", vc->err); b200_vc_close(vc); return NULL; } while (0)
    vc->theta = 1.1e-3;
    vc->flag = CF_ACGT_ONLY;
    vc->output_type = 'v';
    const char *samples_fname = NULL, *ploidy_alias = NULL, *ploidy_fname = NULL, *groups = NULL, *targets = NULL, *gvcf_arg = NULL;
    int samples_is_file = 0, mcall = 0;
    for (int i=0; i<argc; i++)
    {
        const char *a = argv[i];
        #define NEED_ARG(dst) do { if ( i+1>=argc ) OPEN_FAIL("Missing argument to %s\n", a); (dst) = argv[++i]; } while (0)
        if ( a[0]!='-' ) OPEN_FAIL("Unexpected argument: %s\n", a);
        if ( a[1]=='-' )
        {
            const char *v = NULL;
            if ( !strcmp(a, "--no-version") ) continue;
            else if ( !strcmp(a, "--multiallelic-caller") ) mcall = 1;
            else if ( !strcmp(a, "--variants-only") ) vc->aux_flag |= CALL_VARONLY;
            else if ( !strcmp(a, "--keep-alts") ) vc->aux_flag |= CALL_KEEPALT;
            else if ( !strcmp(a, "--insert-missed") ) vc->flag |= CF_INS_MISSED;
            else if ( !strcmp(a, "--keep-masked-refs") ) vc->flag &= ~CF_ACGT_ONLY;
            else if ( !strcmp(a, "--skip-Ns") ) vc->flag |= CF_ACGT_ONLY;
            else if ( !strcmp(a, "--ploidy") ) NEED_ARG(ploidy_alias);
            else if ( !strcmp(a, "--ploidy-file") ) NEED_ARG(ploidy_fname);
            else if ( !strcmp(a, "--group-samples") ) NEED_ARG(groups);
            else if ( !strcmp(a, "--group-samples-tag") ) { NEED_ARG(v); vc->grp_tag = xstrdup(v); }
            else if ( !strcmp(a, "--samples-file") ) { NEED_ARG(samples_fname); samples_is_file = 1; }
            else if ( !strcmp(a, "--samples") ) { NEED_ARG(samples_fname); samples_is_file = 0; }
            else if ( !strcmp(a, "--targets-file") ) NEED_ARG(targets);
            else if ( !strcmp(a, "--gvcf") ) NEED_ARG(gvcf_arg);
            else if ( !strcmp(a, "--annotate") || !strcmp(a, "--format-fields") )
            {
                NEED_ARG(v);
                uint32_t t = parse_output_tags(v);
                if ( t==~0u ) OPEN_FAIL("Could not parse \"%s\"\n", v);
                vc->output_tags |= t;
            }
            else if ( !strcmp(a, "--prior") ) { NEED_ARG(v); char *e; vc->theta = strtod(v, &e); if ( *e ) OPEN_FAIL("Could not parse, expected float argument: -P %s\n", v); }
            else if ( !strcmp(a, "--prior-freqs") )
            {
                NEED_ARG(v);
                const char *c = strchr(v, ',');
                if ( !c ) OPEN_FAIL("Expected two tags with -F (e.g. AN,AC), got \"%s\"\n", v);
                vc->prior_AN = xstrdup(v); vc->prior_AN[c-v] = 0; vc->prior_AC = xstrdup(c+1);
            }
            else if ( !strcmp(a, "--constrain") )
            {
                NEED_ARG(v);
                if ( !strcasecmp(v, "alleles") ) vc->aux_flag |= CALL_CONSTR_ALLELES;
                else OPEN_FAIL("Unsupported argument to -C: \"%s\"\n", v);
            }
            else if ( !strcmp(a, "--skip-variants") )
            {
                NEED_ARG(v);
                if ( !strcasecmp(v, "snps") ) vc->flag |= CF_INDEL_ONLY;
                else if ( !strcasecmp(v, "indels") ) vc->flag |= CF_NO_INDEL;
                else OPEN_FAIL("Unknown skip category \"%s\" (-S argument must be \"snps\" or \"indels\")\n", v);
            }
            else if ( !strcmp(a, "--output-type") ) { NEED_ARG(v); if ( !strchr("vzbu", v[0]) || !v[0] ) OPEN_FAIL("The output type \"%s\" not recognised\n", v); vc->output_type = v[0]; }
            else OPEN_FAIL("Unsupported option: %s\n", a);
            continue;
        }
        /* clustered short options, an argument either attached or next (getopt) */
        for (const char *c = a+1; *c; c++)
        {
            const char *v = NULL;
            int takes = strchr("oOsStTVCPfagFG", *c) != NULL;
            if ( takes )
            {
                if ( c[1] ) v = c+1;
                else NEED_ARG(v);
            }
            switch ( *c )
            {
                case 'm': mcall = 1; break;
                case 'v': vc->aux_flag |= CALL_VARONLY; break;
                case 'A': vc->aux_flag |= CALL_KEEPALT; break;
                case 'i': vc->flag |= CF_INS_MISSED; break;
                case 'M': vc->flag &= ~CF_ACGT_ONLY; break;
                case 'N': vc->flag |= CF_ACGT_ONLY; break;
                case 'G': groups = v; break;
                case 'f': case 'a':
                {
                    uint32_t t = parse_output_tags(v);
                    if ( t==~0u ) OPEN_FAIL("Could not parse \"%s\"\n", v);
                    vc->output_tags |= t;
                    break;
                }
                case 'F':
                {
                    const char *cm = strchr(v, ',');
                    if ( !cm ) OPEN_FAIL("Expected two tags with -F (e.g. AN,AC), got \"%s\"\n", v);
                    vc->prior_AN = xstrdup(v); vc->prior_AN[cm-v] = 0; vc->prior_AC = xstrdup(cm+1);
                    break;
                }
                case 'g': gvcf_arg = v; break;
                case 'O': if ( !strchr("vzbu", v[0]) || !v[0] ) OPEN_FAIL("The output type \"%s\" not recognised\n", v); vc->output_type = v[0]; break;
                case 'C':
                    if ( !strcasecmp(v, "alleles") ) vc->aux_flag |= CALL_CONSTR_ALLELES;
                    else OPEN_FAIL("Unsupported argument to -C: \"%s\"\n", v);
                    break;
                case 'V':
                    if ( !strcasecmp(v, "snps") ) vc->flag |= CF_INDEL_ONLY;
                    else if ( !strcasecmp(v, "indels") ) vc->flag |= CF_NO_INDEL;
                    else OPEN_FAIL("Unknown skip category \"%s\" (-S argument must be \"snps\" or \"indels\")\n", v);
                    break;
                case 'P': { char *e; vc->theta = strtod(v, &e); if ( *e ) OPEN_FAIL("Could not parse, expected float argument: -P %s\n", v); break; }
                case 's': samples_fname = v; samples_is_file = 0; break;
                case 'S': samples_fname = v; samples_is_file = 1; break;
                case 'T': targets = v; break;
                case 'c': OPEN_FAIL("The consensus caller (-c) is not part of this path\n");
                default: OPEN_FAIL("Unsupported option: -%c\n", *c);
            }
            if ( takes ) break;
        }
    }
    /* sanity checks of vcfcall.c:1073-1087 */
    if ( !mcall ) OPEN_FAIL("Expected -m option\n");
    if ( (vc->aux_flag & CALL_CONSTR_ALLELES) && !targets ) OPEN_FAIL("Expected -t or -T with \"-C alleles\"\n");
    if ( (vc->flag & CF_INS_MISSED) && !(vc->aux_flag & CALL_CONSTR_ALLELES) ) OPEN_FAIL("The -i option requires -C alleles\n");
    if ( (vc->aux_flag & CALL_VARONLY) && gvcf_arg ) OPEN_FAIL("The two options cannot be combined: --variants-only and --gvcf\n");

    if ( ploidy_fname )
    {
        char *txt = read_file(ploidy_fname, NULL);
        if ( !txt ) OPEN_FAIL("Could not read the ploidy file: %s\n", ploidy_fname);
        vc->ploidy = b200_ploidy_init_string(txt, 2);
        free(txt);
    }
    else if ( ploidy_alias ) vc->ploidy = b200_ploidy_init_alias(ploidy_alias);
    else vc->ploidy = b200_ploidy_init_string("* * * 0 0\n* * * 1 1\n* * * 2 2\n", 2);
    if ( !vc->ploidy ) OPEN_FAIL("Could not initialize ploidy\n");

    /* init_data (vcfcall.c:608-712) */
    vc->text = vcf_text; vc->len = len;
    vc->hdr = b200_vhdr_parse(vcf_text, len, &vc->off);
    if ( !vc->hdr ) OPEN_FAIL("Failed to read the VCF header\n");
    {
        /* [htslib] every header carries the PASS filter: bcf_hdr_parse puts it behind ##fileformat when the input has none */
        int have = 0;
        for (int i=0; i<vc->hdr->nlines; i++) if ( !strncmp(vc->hdr->lines[i], "##FILTER=<ID=PASS,", 18) ) have = 1;
        if ( !have )
        {
            b200_vhdr_append(vc->hdr, "##FILTER=<ID=PASS,Description=\"All filters passed\">");
            int at = (vc->hdr->nlines>1 && !strncmp(vc->hdr->lines[0], "##fileformat=", 13)) ? 1 : 0;
            char *l = vc->hdr->lines[vc->hdr->nlines-1];
            memmove(&vc->hdr->lines[at+1], &vc->hdr->lines[at], sizeof(char*)*(vc->hdr->nlines-1-at));
            vc->hdr->lines[at] = l;
        }
    }
    if ( targets && tgt_load(vc, targets, (vc->aux_flag & CALL_CONSTR_ALLELES) ? 1 : 0) )
    {
        if ( err && errlen ) snprintf(err, errlen, "%s", vc->err);
        b200_vc_close(vc); return NULL;
    }
    int nhdr = vc->hdr->nsamples;
    int *samples_map = (int*) malloc(sizeof(int)*(nhdr ? nhdr : 1));
    vc->sample2sex = (int*) malloc(sizeof(int)*(nhdr ? nhdr : 1));
    int nsel = nhdr, subset = 0;
    if ( samples_fname )
    {
        char *txt = NULL;
        if ( samples_is_file ) txt = read_file(samples_fname, NULL);
        else
        {
            txt = xstrdup(samples_fname);
            for (char *p = txt; p && *p; p++) if ( *p==',' ) *p = '\n';
        }
        if ( !txt ) { free(samples_map); OPEN_FAIL("Could not read the file: %s\n", samples_fname); }
        int nwarn = 0; char e2[256] = "";
        int rc = b200_samples_parse(txt, (const char *const*)vc->hdr->samples, nhdr, vc->ploidy, samples_map, vc->sample2sex, &nsel, &nwarn, e2, sizeof e2);
        free(txt);
        if ( rc<0 ) { free(samples_map); OPEN_FAIL("%s\n", e2); }
        if ( !nsel ) { free(samples_map); OPEN_FAIL("No matching sample found\n"); }
        for (int i=0; i<nsel; i++) if ( samples_map[i]!=i ) subset = 1;
        if ( nsel!=nhdr ) subset = 1;
    }
    else b200_samples_default(vc->ploidy, nhdr, samples_map, vc->sample2sex);
    vc->nsex = b200_ploidy_nsex(vc->ploidy);
    vc->sex2ploidy_prev = (int*) calloc(vc->nsex ? vc->nsex : 1, sizeof(int));
    vc->nsmpl = nsel;
    vc->ploidy_vec = (uint8_t*) malloc(nsel ? nsel : 1);
    for (int i=0; i<nsel; i++) vc->ploidy_vec[i] = (uint8_t) b200_ploidy_max(vc->ploidy);
    for (int i=0; i<vc->nsex; i++) vc->sex2ploidy_prev[i] = b200_ploidy_max(vc->ploidy);
    for (int i=0; i<nsel; i++) if ( vc->sample2sex[i] >= vc->nsex ) vc->sample2sex[i] = vc->nsex - 1;

    if ( gvcf_arg )     /* gvcf_init + the FORMAT/DP check + gvcf_update_header, vcfcall.c:660-665 */
    {
        gvcf_t *g = vc->gvcf = (gvcf_t*) calloc(1, sizeof(gvcf_t));
        int n = 1;
        for (const char *s = gvcf_arg; *s; s++) if ( *s==',' ) n++;
        g->dp_range = (int*) malloc(sizeof(int)*n);
        const char *ss = gvcf_arg;
        while ( *ss )
        {
            char *se;
            g->dp_range[g->ndp_range++] = (int) strtol(ss, &se, 10);
            if ( se==ss ) { free(samples_map); OPEN_FAIL("Could not parse: --gvcf %s\n", gvcf_arg); }
            if ( *se==',' && se[1] ) { ss = se+1; continue; }
            else if ( !*se ) break;
            free(samples_map); OPEN_FAIL("Could not parse: --gvcf %s\n", gvcf_arg);
        }
        if ( !b200_vhdr_def(vc->hdr, 1, "DP") ) { free(samples_map); OPEN_FAIL("--gvcf output mode requires FORMAT/DP tag, which is not present in the input header\n"); }
        b200_vhdr_append(vc->hdr, "##INFO=<ID=END,Number=1,Type=Integer,Description=\"End position of the variant described in this record\">");
        b200_vhdr_append(vc->hdr, "##INFO=<ID=MinDP,Number=1,Type=Integer,Description=\"Minimum per-sample depth in this gVCF block\">");
    }
    if ( subset && b200_vhdr_subset(vc->hdr, nsel, samples_map) ) { free(samples_map); OPEN_FAIL("Error occurred while subsetting samples\n"); }
    free(samples_map);

    /* mcall_init (mcall.c:361-394): sample groups, header lines */
    if ( groups )
    {
        if ( vc->grp_tag ) { if ( !b200_vhdr_def(vc->hdr, 1, vc->grp_tag) ) OPEN_FAIL("No such FORMAT tag \"%s\"\n", vc->grp_tag); }
        else if ( b200_vhdr_def(vc->hdr, 1, "QS") ) vc->grp_tag = xstrdup("QS");
        else if ( b200_vhdr_def(vc->hdr, 1, "AD") ) vc->grp_tag = xstrdup("AD");
        else OPEN_FAIL("Error: neither \"AD\" nor \"QS\" FORMAT tag exists and no alternative given with -G\n");
        char *txt = strcmp(groups, "-") ? read_file(groups, NULL) : xstrdup("-");
        if ( !txt ) OPEN_FAIL("Could not read the file: %s\n", groups);
        vc->grp_off = (uint32_t*) calloc(vc->nsmpl+2, sizeof(uint32_t));
        vc->grp_smpl = (uint32_t*) calloc(vc->nsmpl+1, sizeof(uint32_t));
        char e2[256] = "";
        int rc = b200_groups_parse(txt, (const char *const*)vc->hdr->samples, vc->nsmpl, vc->grp_off, vc->grp_smpl, &vc->ngroups, e2, sizeof e2);
        free(txt);
        if ( rc<0 ) OPEN_FAIL("%s\n", e2);
        vc->grouped = 1;
    }
    b200_vhdr_append(vc->hdr, "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">");
    if ( vc->output_tags & CALL_FMT_GQ ) b200_vhdr_append(vc->hdr, "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Phred-scaled Genotype Quality\">");
    if ( vc->output_tags & CALL_FMT_GP ) b200_vhdr_append(vc->hdr, "##FORMAT=<ID=GP,Number=G,Type=Float,Description=\"Genotype posterior probabilities in the range 0 to 1\">");
    b200_vhdr_append(vc->hdr, "##INFO=<ID=AC,Number=A,Type=Integer,Description=\"Allele count in genotypes for each ALT allele, in the same order as listed\">");
    b200_vhdr_append(vc->hdr, "##INFO=<ID=AN,Number=1,Type=Integer,Description=\"Total number of alleles in called genotypes\">");
    b200_vhdr_append(vc->hdr, "##INFO=<ID=DP4,Number=4,Type=Integer,Description=\"Number of high-quality ref-forward , ref-reverse, alt-forward and alt-reverse bases\">");
    b200_vhdr_append(vc->hdr, "##INFO=<ID=MQ,Number=1,Type=Integer,Description=\"Average mapping quality\">");
    if ( vc->output_tags & CALL_FMT_PV4 ) b200_vhdr_append(vc->hdr, "##INFO=<ID=PV4,Number=4,Type=Float,Description=\"P-values for strand bias, baseQ bias, mapQ bias and tail distance bias\">");
    /* the likelihood code needs the per-record definitions of QS / I16 until here; the output header does not carry them (vcfcall.c:703-704) */
    b200_vhdr_remove(vc->hdr, 0, "QS");
    b200_vhdr_remove(vc->hdr, 0, "I16");
    b200_vhdr_format(vc->hdr, &vc->out);
    #undef NEED_ARG
    #undef OPEN_FAIL
    return vc;
}

static void vcrec_free(b200_vcrec_t *r)
{
    if ( !r ) return;
    b200_vrec_destroy(r->rec);
    free(r->pre); free(r->PLs); free(r->QS); free(r->ADs); free(r->prior_ac); free(r->ploidy); free(r);
}
void b200_vc_close(b200_vc_t *vc)
{
    if ( !vc ) return;
    b200_vhdr_destroy(vc->hdr);
    b200_ploidy_destroy(vc->ploidy);
    free(vc->sample2sex); free(vc->sex2ploidy_prev); free(vc->ploidy_vec); free(vc->grp_off); free(vc->grp_smpl);
    free(vc->prior_AN); free(vc->prior_AC); free(vc->grp_tag); free(vc->prev_chr);
    for (int i=0; i<vc->ntgt; i++) { for (int j=0; j<vc->tgt[i].n; j++) free(vc->tgt[i].allele[j]); free(vc->tgt[i].allele); free(vc->tgt[i].chr); }
    for (int i=0; i<vc->ntgt_chr; i++) free(vc->tgt_chr[i]);
    free(vc->tgt); free(vc->tgt_chr);
    for (int i=0; i<vc->nbuf; i++) b200_vrec_destroy(vc->buf[i]);
    free(vc->buf);
    if ( vc->gvcf )
    {
        gvcf_t *g = vc->gvcf;
        free(g->dp_range); free(g->dp); free(g->pl); free(g->tmp); free(g->gts); free(g->chr); free(g->als.s); free(g);
    }
    free(vc->out.s);
    free(vc);
}
void b200_vc_call_params(const b200_vc_t *vc, b200_call_t *call)
{
    call->nsmpl = vc->nsmpl;
    call->flag = vc->aux_flag & (CALL_KEEPALT|CALL_VARONLY);
    call->output_tags = vc->output_tags & (CALL_FMT_GQ|CALL_FMT_GP);
    call->theta = vc->theta;
    call->ploidy = vc->ploidy_vec;
    call->unseen = 0;
    call->nsmpl_grp = vc->grouped ? vc->ngroups : 1;
    call->grp_off = vc->grouped ? vc->grp_off : NULL;
    call->grp_smpl = vc->grouped ? vc->grp_smpl : NULL;
    call->use_prior = vc->prior_AN ? 1 : 0;
}
const uint8_t *b200_vc_ploidy(const b200_vc_t *vc) { return vc->ploidy_vec; }
int b200_vc_unseen(const b200_vc_t *vc) { return vc->unseen; }
int b200_vc_output_type(const b200_vc_t *vc) { return vc->output_type; }
const char *b200_vc_output(const b200_vc_t *vc, size_t *len) { if ( len ) *len = vc->out.l; return vc->out.s ? vc->out.s : ""; }
void b200_vc_output_clear(b200_vc_t *vc) { vc->out.l = 0; if ( vc->out.s ) vc->out.s[0] = 0; }

/* ---- reading (next_line, vcfcall.c:471-606) --------------------------------------------------------- */
static b200_vrec_t *read_rec(b200_vc_t *vc)     /* next input record at a targeted position, NULL at the end */
{
    while ( vc->off < vc->len )
    {
        const char *p = vc->text + vc->off;
        const char *e = (const char*) memchr(p, '\n', vc->len - vc->off);
        size_t ll = e ? (size_t)(e-p) : vc->len - vc->off;
        vc->off += ll + (e ? 1 : 0);
        if ( !ll || (ll==1 && p[0]=='\r') ) continue;
        b200_vrec_t *rec = b200_vrec_parse(vc->hdr, p, ll);
        if ( !rec ) { vc_fail(vc, "Error: could not parse the input VCF\n"); return NULL; }
        if ( vc->tgt && !tgt_has_pos(vc, rec->chrom, (uint32_t)rec->pos) ) { b200_vrec_destroy(rec); continue; }   /* exact position, not an interval overlap */
        return rec;
    }
    vc->eof = 1;
    return NULL;
}
static int same_pos(const b200_vrec_t *a, const b200_vrec_t *b) { return a->pos==b->pos && !strcmp(a->chrom, b->chrom); }
static b200_vrec_t *next_line(b200_vc_t *vc)
{
    if ( !(vc->aux_flag & CALL_CONSTR_ALLELES) ) return read_rec(vc);

    /* -C alleles: fill the buffer with the lines of one position, then pair the VCF line and the target line with the
       best matching combination of alleles, same type (SNP vs indel) first */
    int full = 1;
    if ( vc->nbuf==0 ) full = 0;
    else if ( vc->nbuf==1 || same_pos(vc->buf[0], vc->buf[vc->nbuf-1]) ) full = 0;
    if ( !full && !vc->eof )
    {
        b200_vrec_t *rec;
        while ( (rec = read_rec(vc)) )
        {
            if ( vc->nbuf==vc->mbuf ) { vc->mbuf = vc->mbuf ? 2*vc->mbuf : 8; vc->buf = (b200_vrec_t**) realloc(vc->buf, sizeof(*vc->buf)*vc->mbuf); }
            vc->buf[vc->nbuf++] = rec;
            if ( !same_pos(vc->buf[0], rec) ) break;
        }
        if ( vc->err[0] ) return NULL;
    }
    if ( !vc->nbuf ) return NULL;
    b200_vrec_t *rec0 = vc->buf[0];
    int n;
    for (n=vc->nbuf; n>1; n--) if ( same_pos(rec0, vc->buf[n-1]) ) break;
    tgt_t *best_als = NULL; int best_n = 0, best_i = 0;
    /*  The reference walks the target lines of the position with ONE iterator that is not rewound between the buffered
     *  records (vcfcall.c:565-596: regitr_copy before the loop over i), so only the first record ever sees them: the
     *  pairing is "first buffered record x its best unused target line", record after record.  Kept as is.  */
    (void)n;
    for (int i=0; i<1; i++)
    {
        b200_vrec_t *rec = vc->buf[i];
        int rec_indel = is_indel_als(rec->n_allele, rec->allele) ? 1 : -1;
        for (int k=0; k<vc->ntgt; k++)
        {
            tgt_t *als = &vc->tgt[k];
            if ( als->pos!=(uint32_t)rec->pos || strcmp(als->chr, rec->chrom) || als->used ) continue;
            int nmatch = 0, ndref;
            char dref[1024];
            if ( als->n && b200_vcmp_set_ref(rec->allele[0], als->allele[0], dref, sizeof dref, &ndref)==0 )
            {
                nmatch++;
                if ( rec->n_allele > 1 && als->n > 1 )
                    for (int j=1; j<als->n; j++)
                        if ( b200_vcmp_find_allele(dref, ndref, (const char *const*)rec->allele+1, rec->n_allele-1, als->allele[j]) >= 0 ) nmatch++;
            }
            int als_indel = is_indel_als(als->n, als->allele) ? 1 : -1;
            nmatch *= rec_indel*als_indel;
            if ( nmatch > best_n || !best_als ) { best_n = nmatch; best_als = als; best_i = i; }
        }
    }
    vc->cur_tgt = best_als;
    if ( best_als ) best_als->used = 1;
    b200_vrec_t *rec = vc->buf[best_i];
    memmove(&vc->buf[best_i], &vc->buf[best_i+1], sizeof(*vc->buf)*(vc->nbuf-best_i-1));
    vc->nbuf--;
    return rec;
}

/* ---- mcall_constrain_alleles (mcall.c:1271-1421) ----------------------------------------------------- */
static inline int alleles2gt(int a, int b) { return a>b ? a*(a+1)/2+b : b*(b+1)/2+a; }
static inline void gt2alleles(int igt, int *a, int *b)
{
    int k = 0, dk = 1;
    while ( k<igt ) { dk++; k += dk; }
    *b = dk - 1; *a = igt - k + *b;
}
static int constrain_alleles(b200_vc_t *vc, b200_vcrec_t *r, int *unseen)
{
    b200_vrec_t *rec = r->rec;
    tgt_t *tgt = vc->cur_tgt;
    if ( tgt->n > 5 ) return vc_fail(vc, "Maximum accepted number of alleles is 5, got %d\n", tgt->n);
    const char *als[8]; int als_map[8];
    int has_new = 0, nals = 1, ndref;
    char dref[1024];
    if ( b200_vcmp_set_ref(rec->allele[0], tgt->allele[0], dref, sizeof dref, &ndref) < 0 )
        return vc_fail(vc, "The reference alleles are not compatible at %s:%lld .. %s vs %s\n", rec->chrom, (long long)rec->pos+1, tgt->allele[0], rec->allele[0]);
    als[0] = tgt->allele[0]; als_map[0] = 0;
    for (int i=1; i<tgt->n; i++)
    {
        als[nals] = tgt->allele[i];
        int j = b200_vcmp_find_allele(dref, ndref, (const char *const*)rec->allele+1, rec->n_allele-1, tgt->allele[i]);
        if ( j+1==*unseen ) return 1;       /* "Fixme? Cannot constrain to ..": the site is skipped (mcall returns -2) */
        if ( j>=0 ) als_map[nals] = j+1;
        else { als_map[nals] = (*unseen)>=0 ? *unseen : rec->n_allele - 1; has_new = 1; }      /* sic: unseen is never negative */
        nals++;
    }
    if ( *unseen ) { als_map[nals] = *unseen; als[nals] = rec->allele[*unseen]; nals++; }
    if ( !has_new && nals==rec->n_allele ) return 0;
    int nals_ori = rec->n_allele;
    /* the strings of als[] may live in the record: copy before the alleles are replaced */
    char *keep[8];
    for (int i=0; i<nals; i++) keep[i] = xstrdup(als[i]);
    b200_vrec_set_alleles(rec, (const char *const*)keep, nals);
    for (int i=0; i<nals; i++) free(keep[i]);

    int pl_map[36], k = 0;
    for (int i=0; i<nals; i++)
        for (int j=0; j<=i; j++) pl_map[k++] = alleles2gt(als_map[i], als_map[j]);
    int npls_new = k, nsmpl = vc->nsmpl;
    int nPLs = b200_vrec_fmt_ints(rec, "PL", &r->PLs, &r->mPLs);
    if ( nPLs > 0 && nsmpl )
    {
        int npls_ori = nPLs / nsmpl;
        int32_t *new_pl = (int32_t*) malloc(sizeof(int32_t)*(size_t)npls_new*nsmpl);
        for (int i=0; i<nsmpl; i++)
        {
            const int32_t *ori = r->PLs + (size_t)i*npls_ori;
            int32_t *dst = new_pl + (size_t)i*npls_new;
            for (k=0; k<npls_new; k++)
            {
                dst[k] = pl_map[k] < npls_ori ? ori[pl_map[k]] : B200_I32_VECTOR_END;
                if ( dst[k]==B200_I32_MISSING && *unseen>=0 )
                {
                    int k_ori, ia, ib;
                    gt2alleles(pl_map[k], &ia, &ib);
                    k_ori = alleles2gt(ia, *unseen);
                    if ( ori[k_ori]==B200_I32_MISSING ) k_ori = alleles2gt(ib, *unseen);
                    if ( ori[k_ori]==B200_I32_MISSING ) k_ori = alleles2gt(*unseen, *unseen);
                    dst[k] = ori[k_ori];
                }
                if ( !k && dst[k]==B200_I32_VECTOR_END ) dst[k] = B200_I32_MISSING;
            }
        }
        b200_vrec_set_fmt_ints(rec, "PL", new_pl, npls_new*nsmpl);
        free(new_pl);
    }
    /* QS */
    {
        int nqs = b200_vrec_info_floats(rec, "QS", &r->QS, &r->mQS);
        float qs[8];
        for (int i=0; i<nals; i++) qs[i] = als_map[i]<nqs ? r->QS[als_map[i]] : 0;
        b200_vrec_set_info_floats(rec, "QS", qs, nals);
    }
    /* Number=R FORMAT tags */
    for (int i=0; i<rec->n_fmt; i++)
    {
        const b200_vdef_t *d = b200_vhdr_def(vc->hdr, 1, rec->fmt[i].key);
        if ( !d || d->vl!=B200_VL_R || d->type!=B200_HT_INT ) continue;
        int32_t *tmp = NULL; int mtmp = 0;
        char *key = xstrdup(rec->fmt[i].key);
        int nret = b200_vrec_fmt_ints(rec, key, &tmp, &mtmp);
        if ( nret>0 )
        {
            int n1 = nret / nsmpl;
            int32_t *neu = (int32_t*) malloc(sizeof(int32_t)*(size_t)nals*nsmpl);
            for (int j=0; j<nsmpl; j++)
                for (k=0; k<nals; k++) neu[(size_t)j*nals+k] = als_map[k]<n1 ? tmp[(size_t)j*n1 + als_map[k]] : B200_I32_VECTOR_END;
            b200_vrec_set_fmt_ints(rec, key, neu, nals*nsmpl);
            free(neu);
        }
        free(tmp); free(key);
    }
    (void)nals_ori;
    if ( *unseen ) *unseen = nals-1;
    return 0;
}

/* ---- gvcf_write (gvcf.c:88-227) ----------------------------------------------------------------------- */
static int gvcf_write(b200_vc_t *vc, b200_vrec_t *rec, int is_ref)     /* returns 1 when rec itself is to be written */
{
    gvcf_t *g = vc->gvcf;
    int nsmpl = vc->nsmpl, can_collapse = is_ref ? 1 : 0, ret;
    int32_t dp_range = 0, min_dp = 0;
    if ( !rec && !g->prev_range ) return 0;
    int needs_flush = can_collapse ? 0 : 1;
    if ( rec && can_collapse )
    {
        ret = b200_vrec_fmt_ints(rec, "DP", &g->tmp, &g->mtmp);
        if ( ret==nsmpl )
        {
            min_dp = g->tmp[0];
            for (int i=1; i<nsmpl; i++) if ( min_dp > g->tmp[i] ) min_dp = g->tmp[i];
            int i;
            for (i=0; i<g->ndp_range; i++) if ( min_dp < g->dp_range[i] ) break;
            dp_range = i;
            if ( !dp_range ) { needs_flush = 1; can_collapse = 0; }
        }
        else needs_flush = 1;
    }
    if ( g->prev_range && g->prev_range!=dp_range ) needs_flush = 1;
    if ( !rec || !g->chr || strcmp(g->chr, rec->chrom) || rec->pos > g->end+1 ) needs_flush = 1;
    if ( g->prev_range && needs_flush )
    {
        if ( rec && g->chr && !strcmp(rec->chrom, g->chr) && rec->pos==g->end ) g->end--;
        g->end++;
        b200_vrec_t *line = b200_vrec_new(nsmpl);
        line->chrom = g->chr; line->pos = g->start;
        {
            /* bcf_update_alleles_str */
            int n = 1; for (char *p = g->als.s; *p; p++) if ( *p==',' ) n++;
            char **a = (char**) malloc(sizeof(char*)*n); char *copy = xstrdup(g->als.s), *p = copy; int k = 0;
            a[k++] = p;
            for (; *p; p++) if ( *p==',' ) { *p = 0; a[k++] = p+1; }
            b200_vrec_set_alleles(line, (const char *const*)a, n);
            free(a); free(copy);
        }
        int32_t end32 = (int32_t)g->end;
        if ( g->start+1 < g->end ) b200_vrec_set_info_ints(line, "END", &end32, 1);
        b200_vrec_set_info_ints(line, "MinDP", &g->min_dp, 1);
        if ( g->ngts>0 ) b200_vrec_set_genotypes(line, g->gts, g->ngts);
        if ( g->npl>0 ) b200_vrec_set_fmt_ints(line, "PL", g->pl, g->npl);
        b200_vrec_set_fmt_ints(line, "DP", g->dp, nsmpl);
        b200_vrec_format(line, &vc->out);
        line->chrom = (char*)".";
        b200_vrec_destroy(line);
        g->prev_range = 0; free(g->chr); g->chr = NULL; g->npl = 0; g->ngts = 0;
        if ( !rec ) return 0;
    }
    if ( can_collapse )
    {
        if ( !g->prev_range )
        {
            if ( g->mdp < nsmpl ) { g->dp = (int32_t*) realloc(g->dp, sizeof(int32_t)*nsmpl); g->mdp = nsmpl; }
            memcpy(g->dp, g->tmp, sizeof(int32_t)*nsmpl);
            g->npl = b200_vrec_fmt_ints(rec, "PL", &g->pl, &g->mpl);
            g->ngts = b200_vrec_fmt_ints(rec, "GT", &g->gts, &g->mgts);
            free(g->chr); g->chr = xstrdup(rec->chrom);
            g->start = rec->pos;
            g->als.l = 0;
            b200_str_puts(&g->als, rec->allele[0]);
            for (int i=1; i<rec->n_allele; i++) { b200_str_putc(&g->als, ','); b200_str_puts(&g->als, rec->allele[i]); }
            g->min_dp = min_dp;
        }
        else
        {
            if ( g->min_dp > min_dp ) g->min_dp = min_dp;
            for (int i=0; i<nsmpl; i++) if ( g->dp[i] > g->tmp[i] ) g->dp[i] = g->tmp[i];
            ret = b200_vrec_fmt_ints(rec, "PL", &g->tmp, &g->mtmp);
            if ( ret>=0 )
            {
                if ( ret!=nsmpl*3 ) return vc_fail(vc, "Unexpected number of PL fields\n");
                for (int i=0; i<nsmpl; i++)
                {
                    if ( g->pl[3*i+1] > g->tmp[3*i+1] ) { g->pl[3*i+1] = g->tmp[3*i+1]; g->pl[3*i+2] = g->tmp[3*i+2]; }
                    else if ( g->pl[3*i+1]==g->tmp[3*i+1] && g->pl[3*i+2] > g->tmp[3*i+2] ) g->pl[3*i+2] = g->tmp[3*i+2];
                }
            }
            else g->npl = 0;
        }
        g->prev_range = dp_range;
        int32_t *endv = NULL; int mend = 0;
        if ( b200_vrec_info_ints(rec, "END", &endv, &mend)==1 ) g->end = endv[0] - 1;
        else g->end = rec->pos;
        free(endv);
        return 0;
    }
    if ( is_ref && min_dp ) b200_vrec_set_info_ints(rec, "MinDP", &min_dp, 1);
    return 1;
}

/* ---- the steps in front of the call ------------------------------------------------------------------ */
static int emit(b200_vc_t *vc, b200_vrec_t *rec, int is_ref)       /* vcfcall.c:1143-1147 */
{
    if ( vc->gvcf )
    {
        int w = gvcf_write(vc, rec, is_ref);
        if ( w<0 ) return -1;
        if ( !w ) return 0;
    }
    return b200_vrec_format(rec, &vc->out);
}

int b200_vc_next(b200_vc_t *vc, b200_vcrec_t **out, b200_rec_t *in)
{
    for (;;)
    {
        vc->cur_tgt = NULL;
        b200_vrec_t *rec = next_line(vc);
        if ( !rec ) return vc->err[0] ? -1 : 0;
        if ( (vc->aux_flag & CALL_CONSTR_ALLELES) && !vc->cur_tgt ) { b200_vrec_destroy(rec); continue; }      /* duplicate position, every target line used up */
        int is_indel = rec_is_snp(rec) ? 0 : 1;
        if ( ((vc->flag & CF_INDEL_ONLY) && !is_indel) || ((vc->flag & CF_NO_INDEL) && is_indel)
             || ((vc->flag & CF_ACGT_ONLY) && (rec->allele[0][0]=='N' || rec->allele[0][0]=='n')) ) { b200_vrec_destroy(rec); continue; }
        int unseen = b200_unseen_allele((const char *const*)rec->allele, rec->n_allele);
        int is_ref = (rec->n_allele==1 || (rec->n_allele==2 && unseen>0)) ? 1 : 0;
        if ( is_ref && (vc->aux_flag & CALL_VARONLY) ) { b200_vrec_destroy(rec); continue; }
        if ( vc->nsex ) b200_set_ploidy(vc->ploidy, rec->chrom, rec->pos, vc->sample2sex, vc->nsmpl, vc->sex2ploidy_prev, vc->ploidy_vec);

        b200_vcrec_t *r = (b200_vcrec_t*) calloc(1, sizeof *r);
        r->rec = rec;
        if ( vc->flag & CF_INS_MISSED )
        {
            b200_str_t pre = {0,0,0};
            tgt_flush(vc, &pre, rec);
            r->pre = pre.s; r->npre = pre.l;
            free(vc->prev_chr); vc->prev_chr = xstrdup(rec->chrom); vc->prev_beg = (uint32_t)rec->pos; vc->have_prev = 1;
        }
        /* ---- mcall(), the part in front of the likelihoods (mcall.c:1430-1543) ---- */
        if ( vc->aux_flag & CALL_CONSTR_ALLELES )
        {
            int rc = constrain_alleles(vc, r, &unseen);
            if ( rc<0 ) { vcrec_free(r); return -1; }
            if ( rc>0 )     /* mcall returned -2: the site is skipped, lines already flushed in front of it stay */
            {
                if ( r->npre ) b200_str_putsn(&vc->out, r->pre, r->npre);
                vcrec_free(r);
                continue;
            }
        }
        vc->unseen = r->unseen = unseen;
        int nals = r->nals_ori = rec->n_allele, nsmpl = vc->nsmpl;
        r->nPLs = b200_vrec_fmt_ints(rec, "PL", &r->PLs, &r->mPLs);
        if ( r->nPLs!=nsmpl*nals*(nals+1)/2 && r->nPLs!=nsmpl*nals )
        {
            vc_fail(vc, "Wrong number of PL fields? nals=%d npl=%d\n", nals, r->nPLs);
            vcrec_free(r); return -1;
        }
        if ( !vc->grouped )
        {
            r->nQS = b200_vrec_info_floats(rec, "QS", &r->QS, &r->mQS);
            if ( r->nQS<=0 ) { vc_fail(vc, "The QS annotation not present at %s:%lld\n", rec->chrom, (long long)rec->pos+1); vcrec_free(r); return -1; }
        }
        else
        {
            r->nADs = b200_vrec_fmt_ints(rec, vc->grp_tag, &r->ADs, &r->mADs);
            if ( r->nADs<1 ) { vc_fail(vc, "Error: FORMAT/%s is required with the -G option, mpileup must be run with \"-a AD\" or \"-a QS\"\n", vc->grp_tag); vcrec_free(r); return -1; }
        }
        int32_t prior_an = B200_I32_MISSING;
        if ( vc->prior_AN )
        {
            int32_t *an = NULL; int man = 0;
            if ( b200_vrec_info_ints(rec, vc->prior_AN, &an, &man)==1 && an[0] > 0 )
            {
                r->n_prior_ac = b200_vrec_info_ints(rec, vc->prior_AC, &r->prior_ac, &r->m_prior_ac);
                if ( r->n_prior_ac==nals-1 ) prior_an = an[0];      /* otherwise the prior is not applied (mcall.c:1510) */
            }
            free(an);
        }
        b200_vrec_set_info_ints(rec, "QS", NULL, 0);        /* mcall.c:1537 */
        if ( nals > 32 )                                    /* mcall.c:1539-1543: left as it is, mcall returns 0 */
        {
            fprintf(stderr, "Too many alleles at %s:%lld, skipping.\n", rec->chrom, (long long)rec->pos+1);
            if ( r->npre ) b200_str_putsn(&vc->out, r->pre, r->npre);
            if ( !(vc->aux_flag & CALL_VARONLY) || vc->gvcf ) { if ( emit(vc, rec, 0)<0 ) { vcrec_free(r); return -1; } }
            vcrec_free(r);
            continue;
        }
        r->ploidy = (uint8_t*) malloc(nsmpl ? nsmpl : 1);
        memcpy(r->ploidy, vc->ploidy_vec, nsmpl);
        memset(in, 0, sizeof *in);
        in->n_allele = nals;
        in->PLs = r->PLs; in->nPLs = r->nPLs;
        in->QS = r->QS; in->nQS = r->nQS > 0 ? r->nQS : 0;
        in->ADs = r->ADs; in->nADs = r->nADs > 0 ? r->nADs : 0;
        in->prior_an = prior_an; in->prior_ac = r->prior_ac; in->n_prior_ac = prior_an==B200_I32_MISSING ? 0 : r->n_prior_ac;
        in->user = r;
        *out = r;
        return 1;
    }
}

/* ---- the steps behind the call (mcall.c:1576-1684, vcfcall.c:1139-1147) -------------------------------- */
int b200_vc_finish(b200_vc_t *vc, b200_vcrec_t *r, const b200_out_t *o)
{
    b200_vrec_t *rec = r->rec;
    int nsmpl = vc->nsmpl, nals_ori = r->nals_ori, ret = o->ret;
    if ( r->npre ) b200_str_putsn(&vc->out, r->pre, r->npre);
    if ( ret<=0 )       /* not a variant under -v (mcall.c:1567, 1618); without -v the caller always returns the allele count */
    {
        if ( !(vc->aux_flag & CALL_VARONLY) ) { vc_fail(vc, "The caller returned %d at %s:%lld\n", ret, rec->chrom, (long long)rec->pos+1); vcrec_free(r); return -1; }
        vcrec_free(r);
        return 0;
    }
    int nals_new = ret;
    if ( o->als_new==1 ) b200_vrec_set_fmt_ints(rec, "PL", NULL, 0);
    else
    {
        if ( !(o->site_flags & MCB_SITE_REF_GT) )
        {
            if ( (vc->output_tags & CALL_FMT_GP) && o->GPs ) b200_vrec_set_fmt_floats(rec, "GP", o->GPs, o->nPLs);
            if ( (vc->output_tags & CALL_FMT_GQ) && o->GQs ) b200_vrec_set_fmt_ints(rec, "GQ", o->GQs, nsmpl);
        }
        b200_vrec_set_fmt_ints(rec, "PL", o->PLs, o->nPLs);
    }
    if ( nals_ori!=nals_new )       /* mcall_trim_and_update_numberR */
    {
        for (int i=0; i<rec->n_info; i++)
        {
            const b200_vdef_t *d = b200_vhdr_def(vc->hdr, 0, rec->info[i].key);
            if ( !d || d->vl!=B200_VL_R || !rec->info[i].val ) continue;
            /* values are 4-byte words either way: the text tokens are moved, which is what the bit copy amounts to */
            char *val = xstrdup(rec->info[i].val);
            int n = 1; for (char *p = val; *p; p++) if ( *p==',' ) n++;
            char **tok = (char**) malloc(sizeof(char*)*n); int k = 0; char *p = val;
            tok[k++] = p;
            for (; *p; p++) if ( *p==',' ) { *p = 0; tok[k++] = p+1; }
            b200_str_t s = {0,0,0};
            if ( nals_new==1 ) b200_str_puts(&s, tok[0]);
            else
            {
                const char *neu[32]; for (int j=0; j<32; j++) neu[j] = ".";
                for (int j=0; j<nals_ori && j<n; j++) { int l = o->als_map[j]; if ( l>=0 ) neu[l] = tok[j]; }
                for (int j=0; j<nals_new; j++) { if ( j ) b200_str_putc(&s, ','); b200_str_puts(&s, neu[j]); }
            }
            char *key = xstrdup(rec->info[i].key);
            b200_vrec_set_info_text(rec, key, s.s);
            free(key); free(s.s); free(tok); free(val);
        }
        for (int i=0; i<rec->n_fmt; i++)
        {
            const b200_vdef_t *d = b200_vhdr_def(vc->hdr, 1, rec->fmt[i].key);
            if ( !d || d->vl!=B200_VL_R || d->type!=B200_HT_INT ) continue;
            int32_t *tmp = NULL; int mtmp = 0;
            char *key = xstrdup(rec->fmt[i].key);
            int nret = b200_vrec_fmt_ints(rec, key, &tmp, &mtmp);
            if ( nret>0 && nret==nals_ori*nsmpl )
            {
                int32_t *neu = (int32_t*) malloc(sizeof(int32_t)*(size_t)nals_new*nsmpl);
                b200_trim_numberR(tmp, neu, nsmpl, nals_ori, nals_new, o->als_map);
                b200_vrec_set_fmt_ints(rec, key, neu, nals_new*nsmpl);
                free(neu);
            }
            free(tmp); free(key);
        }
    }
    rec->qual = o->qual;
    if ( nals_new>1 ) b200_vrec_set_info_ints(rec, "AC", o->ac+1, nals_new-1);
    { int32_t an = o->an; b200_vrec_set_info_ints(rec, "AN", &an, 1); }
    {
        const char *als[32];
        for (int i=0; i<nals_ori && i<32; i++) if ( o->als_map[i]>=0 ) als[o->als_map[i]] = rec->allele[i];
        char *keep[32];
        for (int i=0; i<nals_new; i++) keep[i] = xstrdup(als[i]);
        b200_vrec_set_alleles(rec, (const char *const*)keep, nals_new);
        for (int i=0; i<nals_new; i++) free(keep[i]);
    }
    b200_vrec_set_genotypes(rec, o->gts, nsmpl*2);
    {
        float *i16 = NULL; int m16 = 0;
        if ( b200_vrec_info_floats(rec, "I16", &i16, &m16)==16 )
        {
            int32_t dp4[4], mq;
            b200_i16_to_dp4_mq(i16, dp4, &mq);
            b200_vrec_set_info_ints(rec, "DP4", dp4, 4);
            b200_vrec_set_info_ints(rec, "MQ", &mq, 1);
            if ( vc->output_tags & CALL_FMT_PV4 )
            {
                float pv[4];
                if ( b200_pv4(i16, pv) ) b200_vrec_set_info_floats(rec, "PV4", pv, 4);
            }
        }
        free(i16);
    }
    b200_vrec_set_info_ints(rec, "I16", NULL, 0);
    int rc = emit(vc, rec, ret==1 ? 1 : 0);
    vcrec_free(r);
    return rc<0 ? -1 : 0;
}

int b200_vc_flush(b200_vc_t *vc)
{
    if ( vc->gvcf && gvcf_write(vc, NULL, 0)<0 ) return -1;
    if ( vc->flag & CF_INS_MISSED ) tgt_flush(vc, &vc->out, NULL);
    return 0;
}

/* ---- the whole command with the device in between ------------------------------------------------------ */
static char g_run_err[512];
static void run_error_handler(const char *msg) { snprintf(g_run_err, sizeof g_run_err, "%s", msg); }

int b200_vcfcall_run(int argc, const char *const *argv, const char *in_path, const char *out_path, int device, char *err, size_t errlen)
{
    size_t len = 0;
    char *text = read_file(in_path, &len);
    if ( !text ) { if ( err ) snprintf(err, errlen, "Failed to read from %s\n", in_path); return -1; }
    if ( len > 2 && (unsigned char)text[0]==31 && (unsigned char)text[1]==139 )     /* BGZF: compressed VCF or BCF (vcfcall.c:624: any htslib-readable input) */
    {
        b200_str_t raw = {0,0,0}, txt = {0,0,0};
        int bad = b200_bgzf_decompress((const uint8_t*)text, len, &raw);
        if ( !bad && raw.l >= 5 && !memcmp(raw.s, "BCF\2\2", 5) ) { bad = b200_bcf_to_vcf_text((const uint8_t*)text, len, &txt); free(raw.s); raw = txt; }
        if ( bad ) { if ( err ) snprintf(err, errlen, "Failed to read from %s: not a BGZF / BCF2 file\n", in_path); free(raw.s); free(text); return -1; }
        free(text); text = raw.s; len = raw.l;
    }
    b200_vc_t *vc = b200_vc_open(argc, argv, text, len, err, errlen);
    if ( !vc ) { free(text); return -1; }
    FILE *fp = strcmp(out_path, "-") ? fopen(out_path, "wb") : stdout;
    if ( !fp ) { if ( err ) snprintf(err, errlen, "Error: cannot write to \"%s\"\n", out_path); b200_vc_close(vc); free(text); return -1; }
    b200_call_t call; memset(&call, 0, sizeof call);
    b200_vc_call_params(vc, &call);
    call.device = device;
    call.max_records = 1024;
    call.max_nals = 32;
    call.async_flush = 1;       /* the next batch is read and unpacked while this one is on the device */
    g_run_err[0] = 0;
    b200_set_error_handler(run_error_handler);
    b200_mcall_init(&call);
    int rc = g_run_err[0] ? -1 : 0, nres;
    b200_vcrec_t *r; b200_rec_t in;
    const int otype = b200_vc_output_type(vc);      /* text streams out batch by batch; the binary forms are encoded from the whole text at the end */
    #define DRAIN(n) do { for (int i_=0; i_<(n) && !rc; i_++) { b200_out_t o_; b200_mcall_result(&call, i_, &o_); if ( b200_vc_finish(vc, (b200_vcrec_t*)o_.user, &o_) ) rc = -1; } \
                          size_t l_; const char *s_ = b200_vc_output(vc, &l_); if ( l_ && otype=='v' ) { fwrite(s_, 1, l_, fp); b200_vc_output_clear(vc); } } while (0)
    while ( !rc && (nres = b200_vc_next(vc, &r, &in)) > 0 )
    {
        call.unseen = (uint8_t) b200_vc_unseen(vc);     /* the ploidy vector is shared storage, rewritten in place by set_ploidy */
        int n = b200_mcall(&call, &in);
        if ( g_run_err[0] ) { rc = -1; break; }
        if ( n>0 ) DRAIN(n);
    }
    if ( !rc && nres<0 ) rc = -1;
    while ( !rc )
    {
        int n = b200_mcall_flush(&call);
        if ( g_run_err[0] ) { rc = -1; break; }
        if ( n<=0 ) break;
        DRAIN(n);
    }
    #undef DRAIN
    if ( !rc && b200_vc_flush(vc) ) rc = -1;
    if ( !rc )
    {
        size_t l; const char *s = b200_vc_output(vc, &l);
        if ( otype=='v' ) { if ( l ) fwrite(s, 1, l, fp); }
        else
        {
            b200_str_t bin = {0,0,0};
            int bad;
            if ( otype=='z' ) { bad = b200_bgzf_compress((const uint8_t*)s, l, 6, &bin); if ( !bad ) bad = b200_bgzf_finish(&bin); }
            else bad = b200_vcf_text_to_bcf(s, l, otype=='b' ? 6 : 0, &bin);
            if ( bad ) { rc = -1; snprintf(g_run_err, sizeof g_run_err, "Error: failed to encode the output\n"); }
            else fwrite(bin.s, 1, bin.l, fp);
            free(bin.s);
        }
    }
    if ( rc && err ) snprintf(err, errlen, "%s", g_run_err[0] ? g_run_err : b200_vc_error(vc));
    b200_mcall_destroy(&call);
    b200_set_error_handler(NULL);
    if ( fp!=stdout ) fclose(fp);
    b200_vc_close(vc);
    free(text);
    return rc;
}
