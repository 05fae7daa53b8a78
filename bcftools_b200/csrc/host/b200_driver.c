/*  b200_driver.c -- ploidy definitions, -G group files and the unseen allele without htslib (include/b200_driver.h).  */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <limits.h>
#include "b200_driver.h"

typedef struct { char *chr; int64_t beg, end; int sex, ploidy; } preg_t;      /* 0-based inclusive */
struct b200_ploidy
{
    int nsex, dflt, min, max, nreg;
    char **id2sex;  int *sex2dflt;
    preg_t *reg;
};

static char *dup_n(const char *s, size_t n) { char *d = (char*) malloc(n+1); if ( d ) { memcpy(d, s, n); d[n] = 0; } return d; }

static int sex_id(const b200_ploidy_t *p, const char *sex)
{
    for (int i=0; i<p->nsex; i++) if ( !strcmp(p->id2sex[i], sex) ) return i;
    return -1;
}
static int sex_new(b200_ploidy_t *p, const char *sex, int dflt)
{
    char **a = (char**) realloc(p->id2sex, sizeof(char*)*(p->nsex+1));
    if ( !a ) return -1;
    p->id2sex = a;
    int *d = (int*) realloc(p->sex2dflt, sizeof(int)*(p->nsex+1));
    if ( !d ) return -1;
    p->sex2dflt = d;
    p->id2sex[p->nsex] = dup_n(sex, strlen(sex));
    p->sex2dflt[p->nsex] = dflt;
    return p->nsex++;
}

/*  one definition line (ploidy_parse, ploidy.c:55-121); 0 ok, 1 skipped (empty / comment), <0 error  */
static int parse_line(b200_ploidy_t *p, const char *ss, const char *end)
{
    const char *f[5]; size_t fl[5]; int nf = 0;
    while ( ss < end && nf < 5 )
    {
        while ( ss < end && isspace((unsigned char)*ss) ) ss++;
        if ( ss >= end ) break;
        const char *se = ss;
        while ( se < end && !isspace((unsigned char)*se) ) se++;
        f[nf] = ss; fl[nf] = (size_t)(se - ss); nf++;
        ss = se;
    }
    if ( !nf || f[0][0]=='#' ) return 1;
    if ( nf < 5 ) return B200_DRV_EPARSE;
    const int is_dflt = fl[0]==1 && f[0][0]=='*';       /* "* * * <sex> <ploidy>" */
    char *sex = dup_n(f[3], fl[3]), *num = dup_n(f[4], fl[4]), *e = NULL;
    if ( !sex || !num ) { free(sex); free(num); return B200_DRV_ENOMEM; }
    const long pl = strtol(num, &e, 10);
    int rc = (e==num) ? B200_DRV_EPARSE : 0;
    int id = -1;
    if ( !rc )
    {
        id = sex_id(p, sex);
        if ( id < 0 ) id = sex_new(p, sex, -1);
        if ( id < 0 ) rc = B200_DRV_ENOMEM;
    }
    if ( !rc )
    {
        if ( p->min<0 || pl < p->min ) p->min = (int)pl;
        if ( p->max<0 || pl > p->max ) p->max = (int)pl;
        /* ploidy.c:114-118 stores the default under the most recently ADDED sex, not under the line's own one: kept */
        if ( is_dflt ) p->sex2dflt[p->nsex-1] = (int)pl;
        else
        {
            char *b = dup_n(f[1], fl[1]), *t = dup_n(f[2], fl[2]), *eb = NULL, *et = NULL;
            const long long beg = b ? strtoll(b, &eb, 10) : 0, to = t ? strtoll(t, &et, 10) : 0;
            if ( !b || !t ) rc = B200_DRV_ENOMEM;
            else if ( eb==b || et==t || beg<1 || to<beg ) rc = B200_DRV_EPARSE;
            else
            {
                preg_t *r = (preg_t*) realloc(p->reg, sizeof(preg_t)*(p->nreg+1));
                if ( !r ) rc = B200_DRV_ENOMEM;
                else
                {
                    p->reg = r;
                    r[p->nreg].chr = dup_n(f[0], fl[0]); r[p->nreg].beg = beg-1; r[p->nreg].end = to-1;
                    r[p->nreg].sex = id; r[p->nreg].ploidy = (int)pl;
                    p->nreg++;
                }
            }
            free(b); free(t);
        }
    }
    free(sex); free(num);
    return rc;
}

b200_ploidy_t *b200_ploidy_init_string(const char *str, int dflt)
{
    b200_ploidy_t *p = (b200_ploidy_t*) calloc(1, sizeof *p);
    if ( !p ) return NULL;
    p->min = p->max = -1;
    const char *ss = str;
    while ( *ss )
    {
        const char *se = ss;
        while ( *se && *se!='\r' && *se!='\n' ) se++;
        if ( parse_line(p, ss, se) < 0 ) { b200_ploidy_destroy(p); return NULL; }
        while ( *se=='\r' || *se=='\n' ) se++;
        ss = se;
    }
    /* _set_defaults (ploidy.c:123-134) */
    const int star = sex_id(p, "*");
    if ( star >= 0 ) dflt = p->sex2dflt[star];
    for (int i=0; i<p->nsex; i++) if ( p->sex2dflt[i]==-1 ) p->sex2dflt[i] = dflt;
    p->dflt = dflt;
    if ( p->min<0 || dflt < p->min ) p->min = dflt;
    if ( p->max<0 || dflt > p->max ) p->max = dflt;
    return p;
}

/*  The assemblies' sex-chromosome layout: first PAR end, second non-PAR start, X length, Y length, MT length.  Males are
 *  haploid on X outside the PARs, on Y and on MT; females have no Y and one MT (vcfcall.c:138-175).  */
typedef struct { const char *alias; long par1_end, nonpar_beg, x_len, y_len, mt_len; } assembly_t;
static const assembly_t assemblies[] =
{
    { "GRCh37", 60000, 2699521, 154931043, 59373566, 16569 },
    { "GRCh38",  9999, 2781480, 155701381, 57227415, 16569 },
};

static int same_nocase(const char *a, const char *b)
{
    while ( *a && *b && tolower((unsigned char)*a)==tolower((unsigned char)*b) ) { a++; b++; }
    return !*a && !*b;
}

b200_ploidy_t *b200_ploidy_init_alias(const char *alias)
{
    char buf[1024];
    for (size_t k=0; k<sizeof assemblies/sizeof assemblies[0]; k++)
    {
        const assembly_t *a = assemblies + k;
        if ( !same_nocase(alias, a->alias) ) continue;
        size_t n = 0;
        const char *pre[2] = { "", "chr" }, *mt[2] = { "MT", "M" };
        for (int v=0; v<2; v++)
            n += (size_t) snprintf(buf+n, sizeof buf - n,
                     "%sX 1 %ld M 1\n%sX %ld %ld M 1\n%sY 1 %ld M 1\n%sY 1 %ld F 0\n%s%s 1 %ld M 1\n%s%s 1 %ld F 1\n",
                     pre[v], a->par1_end, pre[v], a->nonpar_beg, a->x_len, pre[v], a->y_len, pre[v], a->y_len,
                     pre[v], mt[v], a->mt_len, pre[v], mt[v], a->mt_len);
        snprintf(buf+n, sizeof buf - n, "* * * M 2\n* * * F 2\n");
        return b200_ploidy_init_string(buf, 2);
    }
    if ( same_nocase(alias, "X") ) return b200_ploidy_init_string("* * * M 1\n* * * F 2\n", 2);
    if ( same_nocase(alias, "Y") ) return b200_ploidy_init_string("* * * M 1\n* * * F 0\n", 2);
    if ( same_nocase(alias, "1") ) return b200_ploidy_init_string("* * * * 1\n", 2);
    return NULL;
}

void b200_ploidy_destroy(b200_ploidy_t *p)
{
    if ( !p ) return;
    for (int i=0; i<p->nsex; i++) free(p->id2sex[i]);
    for (int i=0; i<p->nreg; i++) free(p->reg[i].chr);
    free(p->id2sex); free(p->sex2dflt); free(p->reg); free(p);
}

int b200_ploidy_add_sex(b200_ploidy_t *p, const char *sex)
{
    const int id = sex_id(p, sex);
    return id >= 0 ? id : sex_new(p, sex, p->dflt);
}
int b200_ploidy_nsex(const b200_ploidy_t *p) { return p->nsex; }
int b200_ploidy_sex2id(const b200_ploidy_t *p, const char *sex) { return sex_id(p, sex); }
const char *b200_ploidy_id2sex(const b200_ploidy_t *p, int id) { return (id<0 || id>=p->nsex) ? NULL : p->id2sex[id]; }
int b200_ploidy_min(const b200_ploidy_t *p) { return p->dflt < p->min ? p->dflt : p->min; }
int b200_ploidy_max(const b200_ploidy_t *p) { return p->dflt > p->max ? p->dflt : p->max; }

int b200_ploidy_query(const b200_ploidy_t *p, const char *seq, int64_t pos, int *sex2ploidy, int *min, int *max)
{
    int hit = 0, mn = INT_MAX, mx = -1;
    for (int i=0; i<p->nreg; i++)
    {
        const preg_t *r = p->reg + i;
        if ( pos < r->beg || pos > r->end || strcmp(r->chr, seq) ) continue;
        if ( !hit && sex2ploidy ) for (int k=0; k<p->nsex; k++) sex2ploidy[k] = p->dflt;
        hit = 1;
        if ( r->ploidy != p->dflt )
        {
            if ( sex2ploidy ) sex2ploidy[r->sex] = r->ploidy;
            if ( mn > r->ploidy ) mn = r->ploidy;
            if ( mx < r->ploidy ) mx = r->ploidy;
        }
    }
    if ( !hit )
    {
        if ( min ) *min = p->dflt;
        if ( max ) *max = p->dflt;
        if ( sex2ploidy ) for (int k=0; k<p->nsex; k++) sex2ploidy[k] = p->sex2dflt[k];
        return 0;
    }
    if ( mx==-1 ) mx = mn = p->dflt;
    if ( min ) *min = mn;
    if ( max ) *max = mx;
    return 1;
}

int b200_set_ploidy(const b200_ploidy_t *p, const char *seq, int64_t pos, const int *sample2sex, int nsmpl,
                    int *sex2ploidy_prev, uint8_t *ploidy)
{
    int cur[64], *s2p = cur, i;
    if ( p->nsex > 64 && !(s2p = (int*) malloc(sizeof(int)*p->nsex)) ) return B200_DRV_ENOMEM;
    b200_ploidy_query(p, seq, pos, s2p, NULL, NULL);
    for (i=0; i<p->nsex; i++) if ( s2p[i]!=sex2ploidy_prev[i] ) break;
    const int changed = i < p->nsex;
    if ( changed )
    {
        for (i=0; i<nsmpl; i++) ploidy[i] = (uint8_t)(sample2sex[i]<0 ? -sample2sex[i] : s2p[sample2sex[i]]);
        memcpy(sex2ploidy_prev, s2p, sizeof(int)*p->nsex);
    }
    if ( s2p != cur ) free(s2p);
    return changed;
}

/* ---- -S samples ----------------------------------------------------------------------------------- */
typedef struct { char **line; int n, m; } lines_t;

static int lines_push(lines_t *l, const char *s, size_t n)
{
    if ( l->n==l->m )
    {
        int m = l->m ? 2*l->m : 16;
        char **a = (char**) realloc(l->line, sizeof(char*)*m);
        if ( !a ) return -1;
        l->line = a; l->m = m;
    }
    if ( !(l->line[l->n] = dup_n(s, n)) ) return -1;
    l->n++;
    return 0;
}
static void lines_free(lines_t *l) { for (int i=0; i<l->n; i++) free(l->line[i]); free(l->line); l->line = NULL; l->n = l->m = 0; }

/*  add_sample (vcfcall.c:114-130): "name sex" unless the name is known already  */
static int ped_add(lines_t *out, const char *name, char sex)
{
    const size_t len = strlen(name);
    for (int i=0; i<out->n; i++) if ( !strncmp(out->line[i], name, len) && out->line[i][len]==' ' && strlen(out->line[i])==len+2 ) return 0;
    char buf[1024];
    if ( len + 3 > sizeof buf ) return -1;
    memcpy(buf, name, len); buf[len] = ' '; buf[len+1] = sex; buf[len+2] = 0;
    return lines_push(out, buf, len+2);
}

/*  parse_ped_samples (vcfcall.c:200-261): 1 = PED (out filled), 0 = not PED, <0 error  */
/*  sample names -> header index without a scan per name (100,000 names would be 1e10 comparisons): an index sorted by name,
 *  ties by position, and a lower-bound search, so that the FIRST header column of a name is the one found  */
typedef struct { const char *s; int i; } name_ix_t;
static int name_ix_cmp(const void *a, const void *b)
{
    const name_ix_t *x = (const name_ix_t*)a, *y = (const name_ix_t*)b;
    const int c = strcmp(x->s, y->s);
    return c ? c : (x->i > y->i) - (x->i < y->i);
}
static name_ix_t *name_index(const char *const *names, int n)
{
    name_ix_t *ix = (name_ix_t*) malloc(sizeof(name_ix_t)*(n ? n : 1));
    if ( !ix ) return NULL;
    for (int i=0; i<n; i++) { ix[i].s = names[i]; ix[i].i = i; }
    qsort(ix, n, sizeof(name_ix_t), name_ix_cmp);
    return ix;
}
static int name_find(const name_ix_t *ix, int n, const char *s, size_t len)
{
    int lo = 0, hi = n;         /* first entry >= (s,len) */
    while ( lo < hi )
    {
        const int mid = lo + (hi - lo)/2;
        int c = strncmp(ix[mid].s, s, len);
        if ( !c && ix[mid].s[len] ) c = 1;
        if ( c < 0 ) lo = mid + 1; else hi = mid;
    }
    if ( lo < n && !strncmp(ix[lo].s, s, len) && !ix[lo].s[len] ) return ix[lo].i;
    return -1;
}

static int ped_parse(const lines_t *in, lines_t *out)
{
    int i;
    for (i=0; i<in->n; i++)
    {
        char *str = dup_n(in->line[i], strlen(in->line[i]));
        if ( !str ) return B200_DRV_ENOMEM;
        /* a PED line has at least six whitespace-separated fields (runs of blanks count once, leading blanks make an empty
           first field): family, sample, father, mother, sex, ...; col[k] = field k+1, cut at its end */
        char *col[5];
        int j = 0;
        for (char *p = str; *p && j<5; )
        {
            p += strcspn(p, " \t\n\v\f\r");            /* end of the current field */
            if ( !*p ) break;
            *p++ = 0;
            p += strspn(p, " \t\n\v\f\r");             /* next field */
            col[j++] = p;
        }
        if ( j!=5 ) { free(str); break; }
        /* columns: family, sample = col[0], father = col[1], mother = col[2], sex = col[3]; each ends at the next separator */
        const char sex = col[3][0]=='1' ? 'M' : 'F';
        int rc = ped_add(out, col[0], sex);
        if ( !rc && strcmp(col[1], "0") && strcmp(col[2], "0") ) { rc = ped_add(out, col[1], 'M'); if ( !rc ) rc = ped_add(out, col[2], 'F'); }
        free(str);
        if ( rc ) return B200_DRV_ENOMEM;
    }
    if ( i!=in->n ) return i>0 ? B200_DRV_EPARSE : 0;      /* "Could not parse samples, not a PED format." / a plain list */
    return 1;
}

void b200_samples_default(const b200_ploidy_t *ploidy, int nhdr, int *samples_map, int *sample2sex)
{
    for (int i=0; i<nhdr; i++) { if ( samples_map ) samples_map[i] = i; sample2sex[i] = ploidy->nsex - 1; }
}

int b200_samples_parse(const char *text, const char *const *hdr_samples, int nhdr, b200_ploidy_t *ploidy,
                       int *samples_map, int *sample2sex, int *nsel, int *nwarn, char *err, size_t errlen)
{
    lines_t in = {0,0,0}, ped = {0,0,0};
    int rc = 0, i;
    const char *ss = text;
    while ( *ss )                       /* hts_readlist: one entry per non-empty line */
    {
        const char *le = ss;
        while ( *le && *le!='\n' && *le!='\r' ) le++;
        if ( le > ss && lines_push(&in, ss, (size_t)(le-ss)) ) { lines_free(&in); return B200_DRV_ENOMEM; }
        while ( *le=='\n' || *le=='\r' ) le++;
        ss = le;
    }
    rc = ped_parse(&in, &ped);
    if ( rc < 0 )
    {
        if ( rc==B200_DRV_EPARSE && err && errlen ) snprintf(err, errlen, "Could not parse samples, not a PED format.");
        lines_free(&in); lines_free(&ped);
        return rc;
    }
    const lines_t *L = rc ? &ped : &in;
    rc = 0;
    int *old2new = (int*) malloc(sizeof(int)*(nhdr ? nhdr : 1));
    name_ix_t *hdr_ix = name_index(hdr_samples, nhdr);
    if ( !old2new || !hdr_ix ) { free(old2new); free(hdr_ix); lines_free(&in); lines_free(&ped); return B200_DRV_ENOMEM; }
    const int dflt_sex = ploidy->nsex - 1;                  /* vcfcall.c:288-289 */
    for (i=0; i<nhdr; i++) { sample2sex[i] = dflt_sex; old2new[i] = -1; }
    int n = 0, warn = 0;
    for (i=0; i<L->n && !rc; i++)
    {
        const char *s0 = L->line[i];
        while ( *s0 && isspace((unsigned char)*s0) ) s0++;
        if ( !*s0 ) { if ( err && errlen ) snprintf(err, errlen, "Could not parse: %s", L->line[i]); rc = B200_DRV_EPARSE; break; }
        if ( *s0=='#' ) continue;
        const char *e0 = s0;
        while ( *e0 && !isspace((unsigned char)*e0) ) e0++;
        const int ismpl = name_find(hdr_ix, nhdr, s0, (size_t)(e0-s0));
        if ( ismpl < 0 ) { warn++; continue; }              /* "Warning: No such sample in the VCF" */
        if ( old2new[ismpl] != -1 ) { warn++; continue; }   /* "Warning: The sample is listed multiple times" */
        const char *s1 = e0;
        while ( *s1 && isspace((unsigned char)*s1) ) s1++;
        char sex[256] = "2";                                /* default ploidy */
        if ( *s1 )
        {
            const char *e1 = s1;
            while ( *e1 && !isspace((unsigned char)*e1) ) e1++;
            size_t l = (size_t)(e1-s1);
            if ( l >= sizeof sex ) l = sizeof sex - 1;
            memcpy(sex, s1, l); sex[l] = 0;
        }
        if ( !sex[1] && (sex[0]=='0' || sex[0]=='1' || sex[0]=='2') ) sample2sex[n] = -(sex[0]-'0');
        else
        {
            const int id = b200_ploidy_add_sex(ploidy, sex);
            if ( id < 0 ) { rc = B200_DRV_ENOMEM; break; }
            sample2sex[n] = id;
        }
        samples_map[n] = ismpl;
        old2new[ismpl] = n;
        n++;
    }
    free(old2new); free(hdr_ix); lines_free(&in); lines_free(&ped);
    if ( rc ) return rc;
    *nsel = n;
    if ( nwarn ) *nwarn = warn;
    return 0;
}

/* ---- -G groups ------------------------------------------------------------------------------------ */
static void set_err(char *err, size_t n, const char *fmt, const char *a)
{
    if ( err && n ) snprintf(err, n, fmt, a);
}

int b200_groups_parse(const char *text, const char *const *samples, int nsmpl, uint32_t *grp_off, uint32_t *grp_smpl,
                      int *ngroups, char *err, size_t errlen)
{
    int i;
    if ( !strcmp(text, "-") )           /* single-sample calling: every sample is its own group (mcall.c:285-296) */
    {
        for (i=0; i<nsmpl; i++) { grp_off[i] = (uint32_t)i; grp_smpl[i] = (uint32_t)i; }
        grp_off[nsmpl] = (uint32_t)nsmpl;
        *ngroups = nsmpl;
        return 0;
    }
    int *smpl2grp = (int*) calloc(nsmpl ? nsmpl : 1, sizeof(int));        /* group + 1, 0 = not listed */
    char **gname = (char**) calloc(nsmpl ? nsmpl : 1, sizeof(char*));
    uint32_t *cnt = (uint32_t*) calloc(nsmpl ? nsmpl : 1, sizeof(uint32_t));
    name_ix_t *smpl_ix = name_index(samples, nsmpl);
    if ( !smpl2grp || !gname || !cnt || !smpl_ix ) { free(smpl2grp); free(gname); free(cnt); free(smpl_ix); return B200_DRV_ENOMEM; }
    int ng = 0, rc = 0;
    const char *ss = text;
    while ( *ss && !rc )
    {
        const char *le = ss;
        while ( *le && *le!='\n' && *le!='\r' ) le++;
        if ( le > ss )                  /* hts_readlist drops empty lines */
        {
            const char *p = ss;
            while ( p<le && !isspace((unsigned char)*p) ) p++;
            const char *name_end = p;
            while ( p<le && isspace((unsigned char)*p) ) p++;
            if ( name_end==le || p==le )
            {
                char *l = dup_n(ss, (size_t)(le-ss));
                set_err(err, errlen, "Could not parse the line, expected a sample name followed by tab and a population name: %s", l ? l : "");
                free(l);
                rc = B200_DRV_EPARSE;
                break;
            }
            /* mcall.c:310-325 leaves ptr on the FIRST character of the population name and keys the hash with ptr+1: groups
               are told apart by the remainder of the line behind that character ("CEU" and "YEU" are one group).  Kept. */
            const char *g = p + 1; size_t gl = (size_t)(le - g);
            const int ismpl = name_find(smpl_ix, nsmpl, ss, (size_t)(name_end-ss));
            if ( ismpl >= 0 )
            {
                if ( smpl2grp[ismpl] )
                {
                    set_err(err, errlen, "Error: the sample \"%s\" is listed twice", samples[ismpl]);
                    rc = B200_DRV_EDUP;
                    break;
                }
                int ig = -1;
                for (i=0; i<ng; i++) if ( strlen(gname[i])==gl && !strncmp(gname[i], g, gl) ) { ig = i; break; }
                if ( ig < 0 ) { gname[ng] = dup_n(g, gl); ig = ng++; }
                cnt[ig]++;
                smpl2grp[ismpl] = ig + 1;
            }
        }
        while ( *le=='\n' || *le=='\r' ) le++;
        ss = le;
    }
    if ( !rc && !ng ) { set_err(err, errlen, "Could not parse the file, no matching samples found%s", ""); rc = B200_DRV_EMISSING; }
    if ( !rc )
        for (i=0; i<nsmpl; i++)
            if ( !smpl2grp[i] ) { set_err(err, errlen, "Error: The sample \"%s\" is not listed", samples[i]); rc = B200_DRV_EMISSING; break; }
    if ( !rc )
    {
        grp_off[0] = 0;
        for (i=0; i<ng; i++) grp_off[i+1] = grp_off[i] + cnt[i];
        for (i=0; i<ng; i++) cnt[i] = 0;
        for (i=0; i<nsmpl; i++) { const int ig = smpl2grp[i] - 1; grp_smpl[grp_off[ig] + cnt[ig]++] = (uint32_t)i; }    /* header order */
        *ngroups = ng;
    }
    for (i=0; i<ng; i++) free(gname[i]);
    free(smpl2grp); free(gname); free(cnt); free(smpl_ix);
    return rc;
}

/* ---- record finaliser pieces ------------------------------------------------------------------------- */
int b200_trim_numberR(const void *src, void *dst, int nvec, int nals_ori, int nals_new, const int8_t *als_map)
{
    const uint32_t *s = (const uint32_t*) src;
    uint32_t *d = (uint32_t*) dst;
    for (int v=0; v<nvec; v++)
        for (int k=0; k<nals_ori; k++)
        {
            const int l = als_map[k];
            if ( l==-1 ) continue;              /* to be dropped */
            d[(size_t)v*nals_new + l] = s[(size_t)v*nals_ori + k];
        }
    return nals_new;
}

void b200_i16_to_dp4_mq(const float *a, int32_t *dp4, int32_t *mq)
{
    for (int i=0; i<4; i++) dp4[i] = (int32_t) a[i];
    const float den = a[0] + a[1] + a[2] + a[3];
    *mq = (int32_t)( (a[8] + a[10]) / den );
}

/* ---- unseen allele ---------------------------------------------------------------------------------- */
int b200_unseen_allele(const char *const *alleles, int n_allele)
{
    for (int i=1; i<n_allele; i++)
    {
        const char *a = alleles[i];
        if ( a[0]=='X' ) return i;                                          /* old X */
        if ( a[0]=='<' && (a[1]=='X' || a[1]=='*') && a[2]=='>' ) return i; /* old <X>, new <*> */
    }
    return 0;
}
