/* config_shim.c -- .conf text reader behind the reference's ConfigVTable (codes/configfile.h), standing in for
 * the flex/bison parser src/modelconfig/{configlex.l,configparser.y,configglue.c,configstore*.c}, whose generated
 * sources are not in the tree and cannot be generated here.  TEST INFRASTRUCTURE (oracle/).
 * The reference's own src/modelconfig/configuration.c is compiled unchanged on top of this. */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <codes/configfile.h>

typedef struct Node {
    char* name;
    int type; /* SE_SECTION / SE_KEY / SE_MULTIKEY */
    char** vals;
    size_t nvals;
    struct Node* children;
    size_t nchildren, cap;
} Node;

static Node* add_child(Node* p, const char* name, int type) {
    if (p->nchildren == p->cap) {
        p->cap = p->cap ? p->cap * 2 : 8;
        p->children = (Node*)realloc(p->children, p->cap * sizeof(Node));
    }
    Node* n = &p->children[p->nchildren++];
    memset(n, 0, sizeof *n);
    n->name = strdup(name);
    n->type = type;
    return n;
}

typedef struct {
    const char* p;
    const char* end;
} Lex;
static void skip(Lex* l) {
    for (;;) {
        while (l->p < l->end && isspace((unsigned char)*l->p)) l->p++;
        if (l->p < l->end && *l->p == '#')
            while (l->p < l->end && *l->p != '\n') l->p++;
        else
            break;
    }
}
static char* token(Lex* l) {
    skip(l);
    if (l->p >= l->end) return NULL;
    if (*l->p == '"') {
        const char* q = ++l->p;
        while (q < l->end && *q != '"') q++;
        char* t = strndup(l->p, (size_t)(q - l->p));
        l->p = q < l->end ? q + 1 : q;
        return t;
    }
    const char* q = l->p;
    while (q < l->end && !isspace((unsigned char)*q) && !strchr("{}=;(),\"#", *q)) q++;
    if (q == l->p) return NULL;
    char* t = strndup(l->p, (size_t)(q - l->p));
    l->p = q;
    return t;
}
static int punct(Lex* l, char c) {
    skip(l);
    if (l->p < l->end && *l->p == c) {
        l->p++;
        return 1;
    }
    return 0;
}
static int parse_body(Lex* l, Node* parent, int top, char** err) {
    for (;;) {
        skip(l);
        if (l->p >= l->end) {
            if (!top) *err = strdup("missing '}'");
            return top ? 0 : -1;
        }
        if (*l->p == '}') {
            if (top) {
                *err = strdup("stray '}'");
                return -1;
            }
            l->p++;
            return 0;
        }
        char* name = token(l);
        if (!name) {
            *err = strdup("syntax error");
            return -1;
        }
        if (punct(l, '{')) {
            Node* n = add_child(parent, name, SE_SECTION);
            free(name);
            size_t idx = parent->nchildren - 1;
            if (parse_body(l, n, 0, err)) return -1;
            (void)idx;
        } else if (punct(l, '=')) {
            Node tmp;
            memset(&tmp, 0, sizeof tmp);
            int multi = 0;
            if (punct(l, '(')) {
                multi = 1;
                for (;;) {
                    if (punct(l, ')')) break;
                    if (punct(l, ',')) continue;
                    char* v = token(l);
                    if (!v) {
                        *err = strdup("bad list");
                        return -1;
                    }
                    tmp.vals = (char**)realloc(tmp.vals, (tmp.nvals + 1) * sizeof(char*));
                    tmp.vals[tmp.nvals++] = v;
                }
            } else {
                char* v = token(l);
                if (!v) {
                    *err = strdup("missing value");
                    return -1;
                }
                tmp.vals = (char**)malloc(sizeof(char*));
                tmp.vals[0] = v;
                tmp.nvals = 1;
            }
            if (!punct(l, ';')) {
                *err = strdup("missing ';'");
                return -1;
            }
            Node* n = add_child(parent, name, multi ? SE_MULTIKEY : SE_KEY);
            free(name);
            n->vals = tmp.vals;
            n->nvals = tmp.nvals;
        } else {
            *err = strdup("expected '{' or '='");
            return -1;
        }
    }
}

static Node* sect(void* handle, SectionHandle s) { return s == ROOT_SECTION ? (Node*)handle : (Node*)s; }
static Node* find(Node* p, const char* name, int want_section) {
    for (size_t i = 0; i < p->nchildren; i++) {
        Node* c = &p->children[i];
        if ((c->type == SE_SECTION) == (want_section != 0) && !strcmp(c->name, name)) return c;
    }
    return NULL;
}
static int v_getKey(void* h, SectionHandle s, const char* key, char* buf, size_t bufsize) {
    Node* n = find(sect(h, s), key, 0);
    if (!n || n->nvals == 0) return -1;
    size_t len = strlen(n->vals[0]);
    if (buf && bufsize) {
        strncpy(buf, n->vals[0], bufsize);
        buf[bufsize - 1] = 0;
    }
    return (int)len;
}
static int v_getMultiKey(void* h, SectionHandle s, const char* key, char*** buf, size_t* e) {
    Node* n = find(sect(h, s), key, 0);
    if (!n) {
        *buf = NULL;
        *e = 0;
        return -1;
    }
    *e = n->nvals;
    *buf = (char**)malloc(sizeof(char*) * (n->nvals ? n->nvals : 1));
    for (size_t i = 0; i < n->nvals; i++) (*buf)[i] = strdup(n->vals[i]);
    return 1;
}
static int v_listSection(void* h, SectionHandle s, SectionEntry* entries, size_t* maxentries) {
    Node* p = sect(h, s);
    size_t n = p->nchildren < *maxentries ? p->nchildren : *maxentries;
    for (size_t i = 0; i < n; i++) {
        entries[i].name = p->children[i].name;
        entries[i].type = (unsigned)p->children[i].type;
    }
    *maxentries = n;
    return (int)p->nchildren;
}
static int v_openSection(void* h, SectionHandle s, const char* name, SectionHandle* out) {
    Node* n = find(sect(h, s), name, 1);
    if (!n) return -1;
    *out = (SectionHandle)n;
    return 1;
}
static int v_getSectionSize(void* h, SectionHandle s, unsigned int* count) {
    *count = (unsigned)sect(h, s)->nchildren;
    return 1;
}
static int v_closeSection(void* h, SectionHandle s) {
    (void)h;
    (void)s;
    return 1;
}
static int v_createSection(void* h, SectionHandle s, const char* name, SectionHandle* out) {
    *out = (SectionHandle)add_child(sect(h, s), name, SE_SECTION);
    return 1;
}
static int v_createKey(void* h, SectionHandle s, const char* key, const char** data, unsigned int count) {
    Node* n = add_child(sect(h, s), key, count > 1 ? SE_MULTIKEY : SE_KEY);
    n->vals = (char**)malloc(sizeof(char*) * (count ? count : 1));
    n->nvals = count;
    for (unsigned i = 0; i < count; i++) n->vals[i] = strdup(data[i]);
    return 1;
}
static void v_free(void* h) { (void)h; }

struct ConfigVTable* txtfile_openStream(FILE* f, char** err) {
    *err = NULL;
    size_t cap = 1 << 16, len = 0;
    char* text = (char*)malloc(cap);
    size_t n;
    while ((n = fread(text + len, 1, cap - len, f)) > 0) {
        len += n;
        if (len == cap) text = (char*)realloc(text, cap *= 2);
    }
    Node* root = (Node*)calloc(1, sizeof(Node));
    root->name = strdup("");
    root->type = SE_SECTION;
    Lex l = {text, text + len};
    if (parse_body(&l, root, 1, err)) {
        free(text);
        return NULL;
    }
    free(text);
    struct ConfigVTable* vt = (struct ConfigVTable*)calloc(1, sizeof *vt);
    vt->getKey = v_getKey;
    vt->getMultiKey = v_getMultiKey;
    vt->listSection = v_listSection;
    vt->openSection = v_openSection;
    vt->getSectionSize = v_getSectionSize;
    vt->closeSection = v_closeSection;
    vt->createSection = v_createSection;
    vt->createKey = v_createKey;
    vt->free = v_free;
    vt->data = root;
    return vt;
}
int txtfile_writeConfig(struct ConfigVTable* h, SectionHandle s, FILE* out, char** err) {
    (void)h; (void)s; (void)out;
    *err = strdup("config dump is not available in the shim");
    return -1;
}
int cf_dump(struct ConfigVTable* cf, SectionHandle h, char** err) {
    (void)cf; (void)h;
    *err = strdup("config dump is not available in the shim");
    return -1;
}
/* YAML front-end (ryml + config_compiler.cxx) is out of scope for the oracle build */
char** yaml_configfile_list_includes(const char* data, size_t len, const char* path, size_t* count) {
    (void)data; (void)len; (void)path;
    *count = 0;
    return NULL;
}
void yaml_configfile_free_includes(char** paths, size_t count) { (void)paths; (void)count; }
struct ConfigVTable* yaml_configfile_load(const char* data, size_t len, const char* const* inc_data, const size_t* inc_lens, size_t n_inc) {
    (void)data; (void)len; (void)inc_data; (void)inc_lens; (void)n_inc;
    fprintf(stderr, "shim: YAML configs are not supported by the oracle build\n");
    exit(2);
}
